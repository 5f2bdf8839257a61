"""flow_net on B200: same module paths as the reference (model.flow_architecture, model.flow_net)."""
