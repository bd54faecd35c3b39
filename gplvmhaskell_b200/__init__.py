"""gplvmhaskell_b200 -- B200 (sm_100a) implementation of the PPCA EM hot path and the GP kernel/Cholesky step
of serokell/GPLVMHaskell behind the reference's own entry points.  All arithmetic runs in the CUDA library
``libgplvm_b200.so`` (C ABI: include/gplvm_b200.h); this package is the thin host-side mirror."""
from ._lib import GplvmError, load as load_library, library_path  # noqa: F401
from .ppca import (PPCA, Left, Right, Session, make_ppca, make_ppca_type_safe, makePPCA, makePPCATypeSafe,  # noqa: F401
                   em_steps_fast, em_steps_missed, emStepsFast, emStepsMissed, check_ppca_input_matrices_prop,
                   with_list_of_indexes, withListOfIndexes)
from .gp import (ard_kernel, kernel_function, kernelFunction, gp_chol_lml, cholesky, tri_solve_lower,  # noqa: F401
                 gp_to_posterior_sample, gpToPosteriorSample)

from .pca import PCA, make_pca, make_pca_type_safe, makePCA  # noqa: F401

__version__ = "0.2.0"
