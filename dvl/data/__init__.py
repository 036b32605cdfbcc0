"""dvl/data of the reference (itm only)."""
