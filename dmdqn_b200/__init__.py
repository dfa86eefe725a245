"""dmdqn_b200: B200-native agent-side hot path of dmdqn."""
