"""zzb200 -- B200-native factorised local-ZigZag event loop behind the ZigZagBoomerang.jl API.

The directory name contains a dot, so it is loaded through ``__graft_entry__.load_package()`` (which registers it
as the module ``zzb200``) rather than by a plain ``import``.
"""
from . import problems  # noqa: F401
from ._capi import BoundError, ZZBError, device_info, init, shutdown  # noqa: F401
from .api import (All, ExtendedForm, SelfMoving, FactBoomerang, FactSampler, FactTrace, GaussianPotential, LocalBound, LogisticSubsampled, Matched, Problem, Run, Trace, ZigZag, discretize, mean, pdmp,  # noqa: F401
                  spdmp, sspdmp, sspdmp2, sspdmp3, sspdmp4, subtrace, trace, cummean, cummean_lists, inclusion_prob)
from .problems import CSC, RectCSC, gmrf_config, grid_precision, logistic_config, replicate_logistic, random_sparse_spd, random_spd, sparse_design  # noqa: F401
from . import multigpu  # noqa: F401,E402
from .multigpu import merge_shards, shard_bounds, spdmp_sharded  # noqa: F401,E402
