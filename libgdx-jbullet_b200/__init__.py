"""libgdx-jbullet_b200 — B200-native collision path (broadphase pairs + narrowphase contacts) behind
libgdx-jbullet's BroadphaseInterface / Dispatcher API.  See DESIGN.md and include/b2c.h.

The directory name carries a hyphen (repo convention), so import it by path:

    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("libgdx_jbullet_b200", ".../libgdx-jbullet_b200/__init__.py",
                                                  submodule_search_locations=[".../libgdx-jbullet_b200"])
    mod = importlib.util.module_from_spec(spec); sys.modules["libgdx_jbullet_b200"] = mod; spec.loader.exec_module(mod)

(__graft_entry__.load_package() does exactly this.)
"""
from . import _lib
from ._lib import B2CError, EXPORTS, LIB_PATH, MANIFOLD_DTYPE, RAW_DTYPE
from .partitioned import PartitionedStepper, partition_check, world_digest
from .world import (ALL_FILTER, DBVT, DEFAULT_FILTER, SAP16, SAP32, STATIC_FILTER, TIGHT, GpuBroadphase, GpuCollisionWorld, GpuDispatcher,
                    GpuPairCache, transforms_to_planes)

__all__ = ["GpuCollisionWorld", "GpuBroadphase", "GpuDispatcher", "GpuPairCache", "B2CError", "TIGHT", "DBVT", "SAP16", "SAP32",
           "DEFAULT_FILTER", "STATIC_FILTER", "ALL_FILTER", "transforms_to_planes", "EXPORTS", "LIB_PATH",
           "MANIFOLD_DTYPE", "RAW_DTYPE", "PartitionedStepper", "partition_check", "world_digest"]
