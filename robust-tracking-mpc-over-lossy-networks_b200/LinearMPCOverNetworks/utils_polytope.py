"""Drop-in module for ``LinearMPCOverNetworks.utils_polytope`` of the reference (re-export)."""
from rtmpc_b200.sets import (calculate_maximum_admissible_output_set, calculate_minimal_robust_positively_invariant_set,  # noqa: F401
                             calculate_RPI, determine_convex_hull, mink_sum, pont_diff, scale, support)
