"""chromevol_enhanced_b200 -- B200-native (sm_100a) chromosome-number likelihood path.

A drop-in for the numpy/scipy hot path of Biols9527/ChromEvol-Enhanced
(`src/ancestral_reconstruction.py --use_chromevol --optimize_params --model_selection`):
rate-matrix assembly, batched FP64 expm, Felsenstein pruning, argmax ancestral states and
forward stochastic mapping, as hand-written CUDA kernels behind the C ABI in
`include/chromevol_b200.h`.  Importing the package does not need a GPU; calling it does, and
there is no CPU fallback.
"""
from .newick import Tree, TreeNode  # noqa: F401
from .flatten import FlatTree  # noqa: F401
from .ancestral_reconstruction import (  # noqa: F401
    MODEL_PARAMETER_NAMES, EVENT_TYPES, ChromEvolutionModel, ChromEvolLikelihoodCalculator,
    ChromEvolOptimizer, StochasticMapper, perform_chromevol_analysis,
    calculate_model_selection_criteria)
from .engine import LikelihoodEngine, params_to_row, rate_matrix, transition_matrices  # noqa: F401
from .pipeline import (  # noqa: F401
    ancestral_reconstruction, quantify_rearrangements_from_tree, write_ancestral_states_report)

__version__ = "0.1.0"
