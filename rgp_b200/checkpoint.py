"""Reader for the reference's checkpoint files.

RGP saves a trained model with GPy's ``model.save(file.h5)`` (examples/walk_run_2_alex.py:565,
svi_experiments/rgp_experiments.py:544,647): an HDF5 file with one dataset ``param_array`` (the flat
optimiser vector) and one dataset per named parameter (``layer_1_rbf_inv_lengthscale``,
``layer_1_inducing_inputs``, ``layer_1_qX_0_mean``, ``layer_1_mlp_layer_1_W`` ...).  h5py / pytables are not
in this image, so this module parses the subset of HDF5 those files use, with numpy only:

  superblock version 0 / 1; "old style" groups (symbol table message -> v1 B-tree of symbol-table nodes +
  local heap for the names); version-1 object headers (with continuation blocks); dataspace message
  v1 / v2; datatype message class 1 (IEEE floating point) and class 0 (fixed point); data layout message
  v3, contiguous, compact or chunked (v1 chunk B-tree), optional deflate + shuffle filters.

``load_checkpoint`` returns {name: ndarray}; ``layer_parameters`` regroups them per layer in the shape
``rgp_b200.layer.DeviceDeepAutoreg`` / ``rgp_b200.backconstraint`` take (lengthscale = 1 / sqrt(inv_l), the
``inv_l=True`` convention of every reference config).
"""
from __future__ import annotations

import re
import struct
import zlib
from typing import Dict, List

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(ValueError):
    pass


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise HDF5FormatError("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise HDF5FormatError("superblock version %d is not supported (files written by GPy's save use 0)" % ver)
        self.so, self.sl = buf[13], buf[14]                       # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise HDF5FormatError("only 8-byte offsets / lengths are supported")
        p = 24 if ver == 0 else 28
        self.base = self.u64(p)
        # root group symbol table entry follows base, free-space, eof, driver addresses
        ste = p + 32
        self.root_header = self.u64(ste + 8)

    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]

    # ---------------------------------------------------------------- object headers (version 1)
    def messages(self, addr: int):
        a = self.base + addr
        if self.b[a] != 1:
            raise HDF5FormatError("object header version %d at %d is not supported" % (self.b[a], addr))
        nmsg = self.u16(a + 2)
        size = self.u32(a + 8)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = self.u16(pos), self.u16(pos + 2), self.b[pos + 4]
                body = pos + 8
                if mtype == 0x10:                                    # continuation
                    blocks.append((self.base + self.u64(body), self.u64(body + 8)))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # ---------------------------------------------------------------- groups (symbol tables)
    def group_entries(self, btree: int, heap: int) -> Dict[str, int]:
        h = self.base + heap
        if self.b[h:h + 4] != b"HEAP":
            raise HDF5FormatError("bad local heap")
        data = self.base + self.u64(h + 24)
        out: Dict[str, int] = {}

        def name_at(off):
            s = data + off
            e = self.b.index(b"\0", s)
            return self.b[s:e].decode("utf-8")

        def walk(node):
            n = self.base + node
            if self.b[n:n + 4] == b"TREE":
                level, used = self.b[n + 5], self.u16(n + 6)
                p = n + 24 + 8                                       # skip siblings, first key
                for _ in range(used):
                    child = self.u64(p)
                    p += 16                                          # child address + next key
                    walk(child)
                _ = level
            elif self.b[n:n + 4] == b"SNOD":
                cnt = self.u16(n + 6)
                p = n + 8
                for _ in range(cnt):
                    out[name_at(self.u64(p))] = self.u64(p + 8)
                    p += 40
            else:
                raise HDF5FormatError("unexpected group node")
        walk(btree)
        return out

    def walk(self, header: int, prefix: str, found: Dict[str, np.ndarray]):
        msgs = self.messages(header)
        stab = [m for m in msgs if m[0] == 0x11]
        if stab:                                                     # a group
            body = stab[0][1]
            for name, child in self.group_entries(self.u64(body), self.u64(body + 8)).items():
                self.walk(child, prefix + name + "/", found)
            return
        kinds = {m[0]: m for m in msgs}
        if 0x01 in kinds and 0x03 in kinds and 0x08 in kinds:        # a dataset
            found[prefix.rstrip("/")] = self.dataset(kinds, msgs)

    # ---------------------------------------------------------------- datasets
    def dataset(self, kinds, msgs) -> np.ndarray:
        # dataspace
        _, body, _ = kinds[0x01]
        ver, rank, flags = self.b[body], self.b[body + 1], self.b[body + 2]
        p = body + (8 if ver == 1 else 4)
        shape = tuple(self.u64(p + 8 * i) for i in range(rank))
        _ = flags
        # datatype
        _, body, _ = kinds[0x03]
        cls = self.b[body] & 0x0F
        bits0 = self.b[body + 1]
        size = self.u32(body + 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 1:
            dt = np.dtype(order + "f%d" % size)
        elif cls == 0:
            dt = np.dtype(order + ("i" if bits0 & 8 else "u") + "%d" % size)
        else:
            raise HDF5FormatError("datatype class %d is not supported" % cls)
        count = int(np.prod(shape)) if shape else 1
        # filters
        filters: List[int] = []
        if 0x0B in kinds:
            _, fb, _ = kinds[0x0B]
            fver, nf = self.b[fb], self.b[fb + 1]
            p = fb + (8 if fver == 1 else 2)
            for _ in range(nf):
                fid = self.u16(p)
                if fver == 1 or fid >= 256:          # id, name length, flags, #client values, name
                    nlen, ncd = self.u16(p + 2), self.u16(p + 6)
                    p += 8 + ((nlen + 7) // 8 * 8 if fver == 1 else nlen)
                else:                                # version 2, library filter: NO name-length field
                    ncd = self.u16(p + 4)            # id, flags, #client values
                    p += 6
                p += 4 * ncd
                if fver == 1 and ncd % 2:
                    p += 4
                filters.append(fid)
        # layout
        _, body, _ = kinds[0x08]
        if self.b[body] != 3:
            raise HDF5FormatError("data layout message version %d is not supported" % self.b[body])
        lclass = self.b[body + 1]
        if lclass == 0:                                              # compact
            n = self.u16(body + 2)
            raw = self.b[body + 4:body + 4 + n]
        elif lclass == 1:                                            # contiguous
            addr, n = self.u64(body + 2), self.u64(body + 10)
            raw = b"\0" * (count * dt.itemsize) if addr == UNDEF else self.b[self.base + addr:self.base + addr + n]
        elif lclass == 2:                                            # chunked
            crank = self.b[body + 2]
            btree = self.u64(body + 3)
            cdims = tuple(self.u32(body + 11 + 4 * i) for i in range(crank - 1))
            return self.chunked(btree, shape, cdims, dt, filters)
        else:
            raise HDF5FormatError("layout class %d" % lclass)
        return np.frombuffer(raw, dtype=dt, count=count).reshape(shape).astype(dt.newbyteorder("="))

    def chunked(self, btree, shape, cdims, dt, filters) -> np.ndarray:
        out = np.zeros(shape, dtype=dt.newbyteorder("="))
        if btree == UNDEF:
            return out
        rank = len(shape)

        def walk(node):
            n = self.base + node
            if self.b[n:n + 4] != b"TREE" or self.b[n + 4] != 1:
                raise HDF5FormatError("bad chunk B-tree")
            level, used = self.b[n + 5], self.u16(n + 6)
            p = n + 24
            keysize = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, _mask = self.u32(p), self.u32(p + 4)
                offs = tuple(self.u64(p + 8 + 8 * i) for i in range(rank))
                child = self.u64(p + keysize)
                p += keysize + 8
                if level > 0:
                    walk(child)
                    continue
                raw = self.b[self.base + child:self.base + child + csize]
                for fid in reversed(filters):
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:                                   # shuffle
                        a = np.frombuffer(raw, dtype=np.uint8)
                        k = dt.itemsize
                        raw = a.reshape(k, -1).T.tobytes() if a.size % k == 0 else raw
                    else:
                        raise HDF5FormatError("filter %d is not supported" % fid)
                chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree)
        return out


def load_checkpoint(path: str) -> Dict[str, np.ndarray]:
    """Every dataset of a GPy ``model.save`` file: {name: ndarray} (group paths joined with '/')."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    found: Dict[str, np.ndarray] = {}
    r.walk(r.root_header, "", found)
    return found


def layer_parameters(ck: Dict[str, np.ndarray]) -> List[Dict[str, np.ndarray]]:
    """Per layer (index = the reference's layer name suffix, 0 = observed layer): kernel / likelihood /
    inducing-input parameters in the naming of ``DeviceDeepAutoreg`` plus whatever latent or
    back-constraint arrays the checkpoint holds for that layer (``qX``, ``init_Xs``, ``X_var``, ``mlp``)."""
    layers: Dict[int, Dict[str, np.ndarray]] = {}
    for name, arr in ck.items():
        m = re.match(r"layer_(\d+)_(.+)$", name.split("/")[-1])
        if not m:
            continue
        i, key = int(m.group(1)), m.group(2)
        d = layers.setdefault(i, {})
        mm = re.match(r"mlp_layer_(\d+)_([Wb])$", key)              # back-constraint network, mlp.py:31-32
        if mm:
            d.setdefault("mlp", {}).setdefault(int(mm.group(1)), {})[mm.group(2)] = np.asarray(arr, dtype=np.float64)
        elif key == "rbf_variance":
            d["variance"] = float(np.ravel(arr)[0])
        elif key == "rbf_inv_lengthscale":
            d["lengthscale"] = 1.0 / np.sqrt(np.asarray(arr, dtype=np.float64).ravel() + 1e-200)   # GPy inv_l convention
        elif key == "rbf_lengthscale":
            d["lengthscale"] = np.asarray(arr, dtype=np.float64).ravel()
        elif key == "inducing_inputs":
            d["Z"] = np.asarray(arr, dtype=np.float64)
        elif key == "Gaussian_noise_variance":
            d["noise_variance"] = float(np.ravel(arr)[0])
        else:
            d[key] = np.asarray(arr)
    for d in layers.values():
        if "mlp" in d:                                              # -> [(W [down, up], b [down]), ...] input to output
            d["mlp"] = [(d["mlp"][k]["W"], d["mlp"][k]["b"]) for k in sorted(d["mlp"])]
    return [layers[i] for i in sorted(layers)]
