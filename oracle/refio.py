"""Reader for the record files oracle/_ref/ref_driver writes (TEST INFRASTRUCTURE ONLY).

Format: 16-byte header (u32 magic 'NTSR', layers, F, up_degree), then tagged arrays:
char name[16]; u32 dtype (0=u32, 1=f32); u64 count; payload. Zero-length tags "batch",
"layer", "end" are structure markers.
"""
import numpy as np


def read_record(path):
    buf = open(path, "rb").read()
    hdr = np.frombuffer(buf, np.uint32, 4, 0)
    assert hdr[0] == 0x4E545352, "bad magic"
    out = dict(layers=int(hdr[1]), F=int(hdr[2]), up_degree=bool(hdr[3]), graph={}, batches=[])
    pos = 16
    cur_batch = None
    cur_layer = None
    while pos < len(buf):
        name = buf[pos:pos + 16].split(b"\0")[0].decode()
        dtype = int(np.frombuffer(buf, np.uint32, 1, pos + 16)[0])
        n = int(np.frombuffer(buf, np.uint64, 1, pos + 20)[0])
        pos += 28
        if name == "end":
            break
        if name == "batch":
            cur_batch = dict(layers=[])
            out["batches"].append(cur_batch)
            cur_layer = None
            continue
        if name == "layer":
            cur_layer = {}
            cur_batch["layers"].append(cur_layer)
            continue
        arr = np.frombuffer(buf, np.uint32 if dtype == 0 else np.float32, n, pos).copy()
        pos += 4 * n
        if cur_batch is None:
            out["graph"][name] = arr
        elif name in ("X0",) or name[0] in "Yd" and name not in ("destination",):
            cur_batch[name] = arr
        else:
            cur_layer[name] = arr
    return out
