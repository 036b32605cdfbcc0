"""dvl/indexer of the reference."""
