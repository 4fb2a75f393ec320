"""Test-support code: Python bindings for the CPU oracle, an independent OBJ/MTL
loader, and scene fixtures.  Test infrastructure only -- never imported by the product."""
