"""mxdetection/models/rpn_heads (/root/reference/README.md:28): the proposal stage (convs are out of scope)."""
from .rpn_head import RPNHead, ProposalConfig  # noqa: F401
from .multi_proposal import MultiProposal, Proposal  # noqa: F401
