"""CPU oracle package — test infrastructure only (see ocflow_oracle.py header)."""
