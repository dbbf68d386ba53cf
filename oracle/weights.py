"""Synthetic weights / poses shared by the oracle-side tools.  The generator itself lives in the
package (nerfmatch_b200/synthetic.py, torch-free) because bench.py's GPU arm needs it too and the
product must not import anything from oracle/."""
from nerfmatch_b200.synthetic import (  # noqa: F401
    K_7SCENES, K_CAMBRIDGE, make_pose, make_state_dict, make_unnorm_scene, uniform01,
)
