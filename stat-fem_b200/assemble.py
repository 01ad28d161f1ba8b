"""``assemble(f)`` convenience function (API of assemble.py in the reference)."""
from .ForcingCovariance import ForcingCovariance
from .InterpolationMatrix import InterpolationMatrix


def assemble(f):
    if isinstance(f, (ForcingCovariance, InterpolationMatrix)):
        f.assemble()
        return f
    raise NotImplementedError("The stat-fem assemble function no longer supports assembling Firedrake objects. "
                              "Use the firedrake assemble function directly")
