"""CPU oracle (test infrastructure only -- see oracle_np.py).  Never imported by the product."""
