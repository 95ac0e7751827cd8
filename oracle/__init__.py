"""CPU oracle of the MPPI rollout path -- TEST INFRASTRUCTURE (see oracle.cpp). PARITY UNPINNED."""
