"""degnorm_b200: B200-native NMF-OA engine, drop-in for the GeneNMFOA path of NUStatBioinfo/DegNorm."""
__version__ = "0.1.0"

from .nmf import GeneNMFOA          # noqa: F401
from .engine import Params, ShardEngine, draw_offsets   # noqa: F401
from .nmf_mpi import run_gene_nmfoa_mpi   # noqa: F401
from .warm_start import load_from_previous   # noqa: F401
from .coverage_merge import merge_coverage, merge_chrom_coverage, merge_overlap_gene_coverage   # noqa: F401
from .gene_filter import filter_genes, DeviceCoverage   # noqa: F401
