"""Stand-in for the absent astropy package (see ../README.md)."""
