"""flax stand-in (the reference imports it and never uses it)."""
