"""CPU oracle -- test infrastructure only (see blr_oracle.py header)."""
