"""How fast does a plain streaming Adam (dmdqn_adam_apply: grid-stride float4 read of grads/theta/m/v, write of theta/m/v) run on
the cfg3 parameter set?  The reference point for K4b's fused Adam epilogue (28 P bytes per network here, 24 P there)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from dmdqn_b200 import _native as N
from dmdqn_b200.group import AgentGroup
n = 256
grp = AgentGroup(n, {"nn_layers": [256, 256], "replay_buffer_size": 300, "batch_size": 256, "precision": "tf32x3"})
rng = np.random.default_rng(0)
for t in range(300):
    s = rng.integers(-1, 20, (n, 89)).astype(np.float32)
    grp.push(s, rng.integers(0, 4, n).astype(np.int32), -rng.random(n) * 100, s, rng.random(n) < 0.1)
grp.learn()
grads = torch.randn_like(grp.theta) * 1e-3
st = torch.cuda.current_stream()
def run():
    N.check(grp.lib.dmdqn_adam_apply(C.byref(grp.dims), C.byref(grp.hp), C.byref(grp.nets), grads.data_ptr(), grp.workspace.data_ptr(),
                                     grp.workspace.numel(), st.cuda_stream))
for _ in range(5): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
p = grp.theta.shape[1]
print(f"adam_apply over {n} networks: {ms*1e3:.1f} us, {28 * p * n / ms / 1e6:.0f} GB/s of 28P bytes")
