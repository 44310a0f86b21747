"""The reference's package name over the B200 engine; see compat/README.md."""
