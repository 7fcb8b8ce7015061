"""B200-native batched solver for the MPC_dist::run_NMPC hot path of HDSRL/SRB-CBF-NMPC.
Import as `srb_cbf_nmpc_b200` (alias module at the repo root; the directory name carries hyphens)."""
from . import capi  # noqa: F401
from .api import BatchSolver, CbfNmpcError, load, default_gains, default_options, LIB_PATH, EXPORTS  # noqa: F401
from .mpc_dist import MPC_dist, MPC_dist_team  # noqa: F401
