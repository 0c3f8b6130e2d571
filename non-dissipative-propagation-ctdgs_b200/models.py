"""`from ctan_b200.models import *` replaces the reference's `from models import *` for the CTAN path
(models/__init__.py:1 -- plus CTAN_tgb, which the reference forgets to export)."""
from .modules import (CTAN, CTAN_tgb, CTANEmbedding, LinkPredictor, SequencePredictor,  # noqa: F401
                      SimpleMemory, TGBLinkPredictor)
from .neighbor_loader import LastNeighborLoader  # noqa: F401

__all__ = ["CTAN", "CTAN_tgb", "CTANEmbedding", "SimpleMemory", "LinkPredictor", "TGBLinkPredictor",
           "SequencePredictor", "LastNeighborLoader"]
