"""dvl/models of the reference (bi_encoder only: the retrieval path)."""
