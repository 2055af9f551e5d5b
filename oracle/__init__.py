"""CPU oracle of the hot path.  TEST INFRASTRUCTURE ONLY -- see oracle/ops.py for the contract and the pinning record."""
