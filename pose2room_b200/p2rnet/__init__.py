"""P2RNet (pose sequence -> 3-D boxes) on the B200 kernels, behind the reference's module / model API.

Class names, constructor signatures `(cfg, optim_spec=None)`, state-dict keys and dtypes are those of
/root/reference/models/p2rnet/modules/* so a reference checkpoint loads unchanged and the reference's
Trainer / Tester (models/p2rnet/training.py, testing.py) drive these modules as they drive their own.
"""
from .registers import METHODS, MODULES, LOSSES  # noqa: F401
from .stgcn import STGCN  # noqa: F401
from .vote_center import CenterVoteModule  # noqa: F401
from .proposal_net import ProposalNet  # noqa: F401
from .loss import BoxNetDetectionLoss, Null  # noqa: F401
from .network import P2RNet  # noqa: F401
