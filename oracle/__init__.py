"""Oracle package: CPU restatements of the reference rspmm algorithm.  TEST INFRASTRUCTURE ONLY -
see rspmm_oracle.py for the contract (who may import this, what it is pinned against)."""
