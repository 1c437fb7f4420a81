"""another_raytracer_b200 — B200-native (sm_100a) wavefront path-tracing backend for the hot path of
blackccpie/another_raytracer, behind the C ABI in include/artb.h. See DESIGN.md."""
from ._abi import LIB_PATH, SYMBOLS, load  # noqa: F401
from .engine import (ArtbError, CameraDesc, Context, Engine, EngineMode, HIT_DTYPE, MATERIAL_NAMES, RenderOptions,  # noqa: F401
                     SHADE_DTYPE, Scene, SceneAlias)

__all__ = ["ArtbError", "CameraDesc", "Context", "Engine", "EngineMode", "HIT_DTYPE", "MATERIAL_NAMES", "RenderOptions",
           "SHADE_DTYPE", "Scene", "SceneAlias", "LIB_PATH", "SYMBOLS", "load"]
