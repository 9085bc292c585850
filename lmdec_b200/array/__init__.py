from .chained import ChainedArray
from .matrix_ops import diag_dot, subspace_to_SVD, subspace_to_V, svd_to_trunc_svd
from .metrics import q_value_converge, rmse_k, subspace_dist
from .operators import DenseOperator, SparseOperator, as_operator
from .random import array_constant_partition, array_split, cumulative_partition
from .scaled_legacy import ArrayMoment, ScaledCenterArray
from .stacked import StackedArray
from .transform import acc_format_svd
