"""Stand-in for the absent novas package (see ../README.md)."""
