"""seeded random QPs over every compiled shape at edge horizons (1, 2, 3, 4, 5, 6, 7, 33, 64) and partial tiles against the oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import python_mpc_b200 as pm
import parity_cases as pc
be = pm.cuda_backend()
for B in (3, 37):
    w = pc.check_random_problems(be, seeds=tuple(range(2, 11)), B=B, horizons=(1, 2, 3, 4, 5, 6, 7, 33, 64))
    print("B=%d ok:" % B, w)
