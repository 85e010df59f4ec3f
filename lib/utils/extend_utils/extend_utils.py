"""`lib.utils.extend_utils.extend_utils` as lib/utils/evaluation_utils.py:6 imports it
(`from lib.utils.extend_utils.extend_utils import uncertainty_pnp, ...`): the uncertainty-driven PnP
served by pvnet_b200's device solver.  The module's other functions (mesh rasterisation, farthest point
sampling, nearest neighbours) are dataset tooling outside the inference hot path and are not provided."""
from pvnet_b200.extend_utils import covariance_to_weights, uncertainty_pnp, uncertainty_pnp_batched  # noqa: F401

__all__ = ["uncertainty_pnp", "uncertainty_pnp_batched", "covariance_to_weights"]
