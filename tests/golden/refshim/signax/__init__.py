"""Stand-in for signax 0.1.1's signature / log / flatten (see ../README.md): plain tensor-algebra definitions
written independently of oracle/tensor_algebra.py."""
