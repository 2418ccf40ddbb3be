"""Oracle: 6-vector pose -> affine transform.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows cv_ops.py:20-76 (``geometric_transform``).
"""
import math

import torch

TWO_PI = 2.0 * math.pi


def pose_to_affine(raw, similarity=False, nonlinear=True, as_matrix=False):
    """raw [..., 6] = (scale_x, scale_y, theta, shear, trans_x, trans_y) -> [..., 6] or [..., 3, 3].

    cv_ops.py:36-38 split; :40-45 non-linearities (sigmoid+1e-2 scales, tanh(5x) translations/shear, theta*2pi);
    :47 |.|+1e-2 when linear; :49 cos/sin; :51-63 the two row layouts; :68-74 homogeneous row.
    """
    sx, sy, theta, shear, tx, ty = (raw[..., i:i + 1] for i in range(6))
    if nonlinear:
        sx = torch.sigmoid(sx) + 1e-2
        sy = torch.sigmoid(sy) + 1e-2
        tx = torch.tanh(tx * 5.)
        ty = torch.tanh(ty * 5.)
        shear = torch.tanh(shear * 5.)
        theta = theta * TWO_PI
    else:
        sx = abs(sx) + 1e-2
        sy = abs(sy) + 1e-2
    c = torch.cos(theta)
    s = torch.sin(theta)
    if similarity:
        rows = [sx * c, -sx * s, tx, sx * s, sx * c, ty]
    else:
        rows = [sx * c + shear * sy * s, -sx * s + shear * sy * c, tx, sy * s, sy * c, ty]
    out = torch.cat(rows, -1)
    if as_matrix:
        out = out.view(*out.shape[:-1], 2, 3)
        bottom = torch.zeros_like(out[..., :1, :])
        bottom[..., 0, 2] = 1
        out = torch.cat([out, bottom], -2)
    return out
