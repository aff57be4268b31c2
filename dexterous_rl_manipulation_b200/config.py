"""Curriculum parameters as the batched env consumes them.

The reference's ``experiments.CurriculumConfig`` (experiments/config.py:17-217) is accepted
unchanged wherever a config is expected -- only its attributes are read (duck typing), so
``CurriculumScheduler`` and the config JSONs keep working.  ``CurriculumConfig`` below is a
field-compatible stand-in for users who do not have the reference on their path; the presets
carry the values of experiments/config.py:174-217.
"""
import json
from dataclasses import asdict, dataclass
from typing import Optional, Sequence, Tuple

from . import _lib

Range = Optional[Tuple[float, float]]

# (size, mass, friction, spawn_distance) of the named presets, experiments/config.py:174-217
_PRESETS = {"easy": (0.08, 0.05, 0.8, 0.10), "medium": (0.05, 0.1, 0.5, 0.15), "hard": (0.03, 0.2, 0.3, 0.20)}


def _maybe_uniform(rng, interval: Range, fixed: float) -> float:
    """One uniform draw when the field is randomised, the fixed value (and NO draw) otherwise --
    the draw-count behaviour the reset sequence of envs/manipulation_env.py:151-153 depends on."""
    return fixed if interval is None else float(rng.uniform(interval[0], interval[1]))


@dataclass
class CurriculumConfig:
    object_size: float = 0.05
    object_size_range: Range = None
    object_mass: float = 0.1
    object_mass_range: Range = None
    friction_coefficient: float = 0.5
    friction_range: Range = None
    spawn_distance: float = 0.15            # carried for schema compatibility; unused by the dynamics
    spawn_distance_range: Range = None      # (experiments/config.py:86 has no caller)
    spawn_x_range: Tuple[float, float] = (-0.1, 0.1)
    spawn_y_range: Tuple[float, float] = (-0.1, 0.1)
    spawn_z_range: Tuple[float, float] = (0.05, 0.2)

    # -- samplers (names and draw order are part of the env contract)
    def get_object_size(self, rng) -> float:
        return _maybe_uniform(rng, self.object_size_range, self.object_size)

    def get_object_mass(self, rng) -> float:
        return _maybe_uniform(rng, self.object_mass_range, self.object_mass)

    def get_friction_coefficient(self, rng) -> float:
        return _maybe_uniform(rng, self.friction_range, self.friction_coefficient)

    def get_spawn_distance(self, rng) -> float:
        return _maybe_uniform(rng, self.spawn_distance_range, self.spawn_distance)

    def get_spawn_position(self, rng):
        return tuple(float(rng.uniform(*axis)) for axis in (self.spawn_x_range, self.spawn_y_range, self.spawn_z_range))

    # -- (de)serialisation compatible with the reference's JSON files (experiments/config_*.json)
    def to_dict(self):
        return asdict(self)

    @classmethod
    def from_dict(cls, d):
        return cls(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items()})

    @classmethod
    def from_json(cls, path):
        with open(path) as fh:
            return cls.from_dict(json.load(fh))

    def to_json(self, path):
        with open(path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=2)

    @classmethod
    def preset(cls, name: str):
        size, mass, friction, distance = _PRESETS[name]
        return cls(object_size=size, object_mass=mass, friction_coefficient=friction, spawn_distance=distance)


for _name in _PRESETS:          # CurriculumConfig.easy() / .medium() / .hard()
    setattr(CurriculumConfig, _name, classmethod(lambda cls, _n=_name: cls.preset(_n)))


def _pair(r):
    return None if r is None else (float(r[0]), float(r[1]))


def group_from_config(cfg, sigma_obs: float = 0.0, sigma_dyn: float = 0.0) -> "_lib.DexsimGroup":
    """One row of the device group table from any CurriculumConfig-like object."""
    g = _lib.DexsimGroup()
    g.size = float(cfg.object_size)
    g.mass = float(cfg.object_mass)
    g.friction = float(cfg.friction_coefficient)
    for prefix, rng in (("size", _pair(getattr(cfg, "object_size_range", None))),
                        ("mass", _pair(getattr(cfg, "object_mass_range", None))),
                        ("fric", _pair(getattr(cfg, "friction_range", None)))):
        if rng is not None:
            setattr(g, prefix + "_lo", rng[0])
            setattr(g, prefix + "_hi", rng[1])
            setattr(g, prefix + "_ranged", 1)
    sx = _pair(getattr(cfg, "spawn_x_range", (-0.1, 0.1)))
    sy = _pair(getattr(cfg, "spawn_y_range", (-0.1, 0.1)))
    sz = _pair(getattr(cfg, "spawn_z_range", (0.05, 0.2)))
    for k, (lo, hi) in enumerate((sx, sy, sz)):
        g.spawn_lo[k] = lo
        g.spawn_hi[k] = hi
    g.sigma_obs = float(sigma_obs)
    g.sigma_dyn = float(sigma_dyn)
    return g


def group_table(configs: Sequence, sigma_obs=0.0, sigma_dyn=0.0):
    """ctypes array of groups; ``sigma_*`` may be scalars or per-group sequences."""
    n = len(configs)
    if not 1 <= n <= _lib.MAX_GROUPS:
        raise ValueError(f"need 1..{_lib.MAX_GROUPS} groups, got {n}")
    so = list(sigma_obs) if hasattr(sigma_obs, "__len__") else [sigma_obs] * n
    sd = list(sigma_dyn) if hasattr(sigma_dyn, "__len__") else [sigma_dyn] * n
    arr = (_lib.DexsimGroup * n)()
    for k, cfg in enumerate(configs):
        arr[k] = group_from_config(cfg, so[k], sd[k])
    return arr
