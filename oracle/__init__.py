"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT) — see oracle/pt_oracle.cpp."""
