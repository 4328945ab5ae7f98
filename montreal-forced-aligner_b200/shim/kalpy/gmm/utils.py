from mfa_b200.kalpy_compat import read_gmm_model, read_topology, read_transition_model, read_tree, write_gmm_model  # noqa: F401
