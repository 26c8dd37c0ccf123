"""Test infrastructure: CPU oracle for cc3d_b200 (never imported by the product package)."""
