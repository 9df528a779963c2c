"""Oracle = test infrastructure. See oracle_np.py for the contract (who may import this)."""
