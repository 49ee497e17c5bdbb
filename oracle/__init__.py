"""CPU oracle of the paillier-halo2 hot path — test infrastructure only (see paillier_oracle.py)."""
