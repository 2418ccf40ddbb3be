"""Named-array files between the tests and the emulation harnesses (tests/emu/emu_io.h is the C++ side)."""
import struct

import numpy as np
import torch


def write_arrays(path, arrays):
    """arrays: {name: torch tensor | numpy array | None}; None entries are skipped (absent = NULL on the C++ side)."""
    items = [(k, v) for k, v in arrays.items() if v is not None]
    with open(path, 'wb') as f:
        f.write(struct.pack('<I', len(items)))
        for name, v in items:
            a = np.ascontiguousarray(v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
            f.write(struct.pack('<I', len(name)))
            f.write(name.encode())
            f.write(struct.pack('<IQ', a.dtype.itemsize, a.size))
            f.write(a.tobytes())


def read_arrays(path, dtypes):
    """-> {name: 1-D torch tensor}; ``dtypes`` maps names to numpy dtypes (default float32)."""
    out = {}
    with open(path, 'rb') as f:
        (n,) = struct.unpack('<I', f.read(4))
        for _ in range(n):
            (ln,) = struct.unpack('<I', f.read(4))
            name = f.read(ln).decode()
            es, cnt = struct.unpack('<IQ', f.read(12))
            dt = np.dtype(dtypes.get(name, np.float32))
            assert dt.itemsize == es, (name, dt, es)
            out[name] = torch.from_numpy(np.frombuffer(f.read(es * cnt), dtype=dt).copy())
    return out
