"""`Rays` container of the reference (datasets/base_datasets.py:13-21), same field order."""
import collections

Rays = collections.namedtuple(
    "Rays", ("origins", "directions", "viewdirs", "radii", "lossmult", "near", "far", "noise_var"))
Rays_keys = Rays._fields


def namedtuple_map(fn, tup):
    """Apply `fn` to each element of `tup` and cast to `tup`'s namedtuple (base_datasets.py:19-21)."""
    return type(tup)(*map(fn, tup))
