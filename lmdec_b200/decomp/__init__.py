from .init_methods import rnormal_start, sub_svd_init, sym_mat_mult_start, v_init
from .iter_methods import PowerMethod, SuccessiveBatchedPowerMethod
from .utils import PowerMethodSummary, get_array_moments, make_snp_array, mask_imputation
