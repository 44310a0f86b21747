"""Import names of Blocks that lvsr configs and scripts reference; see compat/README.md."""
