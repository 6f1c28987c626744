"""strainflair_b200 -- B200-native (sm_100a) implementation of StrainFLAIR's query-side abundance engine
(the reference's json2csv + compute_strains_abundance path).  See DESIGN.md."""
__version__ = "0.1.0"

from .schema import Graph, Reads  # noqa: F401
