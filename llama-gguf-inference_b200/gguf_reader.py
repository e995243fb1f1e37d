"""GGUF v3 container reader and writer (host side, no third-party dependency).

Replaces the model-file loader of the reference's backend (`llama-server -m <path>`,
/root/reference/scripts/start.sh:473-475; model resolution start.sh:309-343).  The on-disk format is the one
documented by gguf-py (site-packages/gguf/gguf_reader.py:132-185 header/alignment, :217-257 values,
:259-287 tensor infos; constants.py:10-12 magic/version/alignment, :4157-4170 value types,
:4059-4093 tensor types):

    u32 magic 'GGUF' | u32 version (2 or 3) | u64 n_tensors | u64 n_kv
    n_kv   x { string key | u32 value_type | value }        string = u64 len + bytes
                                                           array  = u32 elem_type | u64 count | elems
    n_tens x { string name | u32 n_dims | u64 dims[n_dims] (ne0 = innermost first) | u32 type | u64 offset }
    pad to general.alignment (default 32) ; tensor data, each tensor at data_start + offset

Only little-endian files are supported (what llama.cpp writes on x86/arm64).
"""
from __future__ import annotations

import mmap
import os
import struct
from dataclasses import dataclass

import numpy as np

GGUF_MAGIC = 0x46554747
DEFAULT_ALIGNMENT = 32

# value types
T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 = range(13)
_SCALAR = {T_U8: "<B", T_I8: "<b", T_U16: "<H", T_I16: "<h", T_U32: "<I", T_I32: "<i", T_F32: "<f",
           T_BOOL: "<?", T_U64: "<Q", T_I64: "<q", T_F64: "<d"}
_NP = {T_U8: np.uint8, T_I8: np.int8, T_U16: np.uint16, T_I16: np.int16, T_U32: np.uint32, T_I32: np.int32,
       T_F32: np.float32, T_BOOL: np.bool_, T_U64: np.uint64, T_I64: np.int64, T_F64: np.float64}

# ggml tensor types this engine understands: id -> (name, block elements, block bytes)
GGML_F32, GGML_F16, GGML_Q8_0, GGML_Q4_K, GGML_Q5_K, GGML_Q6_K, GGML_BF16 = 0, 1, 8, 12, 13, 14, 30
GGML_Q4_0, GGML_Q5_0 = 2, 6       # legacy 32-element blocks: carried as Q8_0 after an exact load-time conversion (model.py)
GGML_TYPES = {
    GGML_F32: ("F32", 1, 4), GGML_F16: ("F16", 1, 2), GGML_BF16: ("BF16", 1, 2),
    GGML_Q8_0: ("Q8_0", 32, 34), GGML_Q4_K: ("Q4_K", 256, 144), GGML_Q5_K: ("Q5_K", 256, 176),
    GGML_Q6_K: ("Q6_K", 256, 210), GGML_Q4_0: ("Q4_0", 32, 18), GGML_Q5_0: ("Q5_0", 32, 22),
}
# every id gguf defines, so that unsupported tensors are named in the error instead of "type 10"
_ALL_TYPE_NAMES = {3: "Q4_1", 7: "Q5_1", 9: "Q8_1", 10: "Q2_K", 11: "Q3_K", 15: "Q8_K",
                   16: "IQ2_XXS", 17: "IQ2_XS", 18: "IQ3_XXS", 19: "IQ1_S", 20: "IQ4_NL", 21: "IQ3_S", 22: "IQ2_S",
                   23: "IQ4_XS", 24: "I8", 25: "I16", 26: "I32", 27: "I64", 28: "F64", 29: "IQ1_M", 34: "TQ1_0",
                   35: "TQ2_0", 39: "MXFP4"}


class GGUFError(ValueError):
    pass


def type_name(t: int) -> str:
    if t in GGML_TYPES:
        return GGML_TYPES[t][0]
    return _ALL_TYPE_NAMES.get(t, f"type{t}")


def row_bytes(t: int, k: int) -> int:
    if t not in GGML_TYPES:
        raise GGUFError(f"unsupported tensor type {type_name(t)}")
    _, be, bb = GGML_TYPES[t]
    if k % be:
        raise GGUFError(f"row length {k} is not a multiple of the {type_name(t)} block size {be}")
    return k // be * bb


@dataclass
class TensorInfo:
    name: str
    ggml_type: int
    ne: tuple          # ne[0] = innermost (K); a 2-D weight [rows, K] has ne = (K, rows)
    offset: int        # absolute file offset
    nbytes: int

    @property
    def n_elements(self) -> int:
        n = 1
        for d in self.ne:
            n *= d
        return n

    @property
    def type_name(self) -> str:
        return type_name(self.ggml_type)


class GGUFFile:
    """Parsed GGUF file: `.meta` (dict key -> python value / numpy array / list of str), `.tensors`
    (name -> TensorInfo), `.data(name)` -> uint8 view of the tensor's bytes (zero-copy on the mmap)."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        if size < 24:
            raise GGUFError(f"{path}: too small to be a GGUF file ({size} bytes)")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        self._buf = memoryview(self._mm)
        self.size = size
        self.meta: dict = {}
        self.meta_types: dict = {}
        self.tensors: dict[str, TensorInfo] = {}
        self._parse()

    # -- low level
    def _need(self, off: int, n: int):
        if off + n > self.size or n < 0:
            raise GGUFError(f"{self.path}: truncated (need {n} bytes at {off}, file has {self.size})")

    def _u32(self, off):
        self._need(off, 4)
        return struct.unpack_from("<I", self._buf, off)[0], off + 4

    def _u64(self, off):
        self._need(off, 8)
        return struct.unpack_from("<Q", self._buf, off)[0], off + 8

    def _str(self, off):
        n, off = self._u64(off)
        self._need(off, n)
        return bytes(self._buf[off:off + n]).decode("utf-8", errors="replace"), off + n

    def _value(self, vt: int, off: int):
        if vt == T_STR:
            return self._str(off)
        if vt in _SCALAR:
            sz = struct.calcsize(_SCALAR[vt])
            self._need(off, sz)
            return struct.unpack_from(_SCALAR[vt], self._buf, off)[0], off + sz
        if vt == T_ARR:
            et, off = self._u32(off)
            cnt, off = self._u64(off)
            if et == T_STR:
                out = []
                for _ in range(cnt):
                    s, off = self._str(off)
                    out.append(s)
                return out, off
            if et in _NP:
                dt = np.dtype(_NP[et]).newbyteorder("<")
                self._need(off, cnt * dt.itemsize)
                arr = np.frombuffer(self._buf, dtype=dt, count=cnt, offset=off)
                return arr, off + cnt * dt.itemsize
            if et == T_ARR:
                out = []
                for _ in range(cnt):
                    v, off = self._value(T_ARR, off)
                    out.append(v)
                return out, off
            raise GGUFError(f"{self.path}: unknown array element type {et}")
        raise GGUFError(f"{self.path}: unknown metadata value type {vt}")

    def _parse(self):
        magic, off = self._u32(0)
        if magic != GGUF_MAGIC:
            raise GGUFError(f"{self.path}: bad magic 0x{magic:08x} (not a GGUF file)")
        self.version, off = self._u32(off)
        if self.version not in (2, 3):
            raise GGUFError(f"{self.path}: unsupported GGUF version {self.version}")
        n_tensors, off = self._u64(off)
        n_kv, off = self._u64(off)
        for _ in range(n_kv):
            key, off = self._str(off)
            vt, off = self._u32(off)
            val, off = self._value(vt, off)
            self.meta[key] = val
            self.meta_types[key] = vt
        self.alignment = int(self.meta.get("general.alignment", DEFAULT_ALIGNMENT))
        if self.alignment <= 0 or self.alignment & (self.alignment - 1):
            raise GGUFError(f"{self.path}: general.alignment {self.alignment} is not a power of two")
        infos = []
        for _ in range(n_tensors):
            name, off = self._str(off)
            nd, off = self._u32(off)
            if nd > 4:
                raise GGUFError(f"{self.path}: tensor {name} has {nd} dims")
            ne = []
            for _ in range(nd):
                d, off = self._u64(off)
                ne.append(int(d))
            tt, off = self._u32(off)
            rel, off = self._u64(off)
            infos.append((name, tuple(ne), tt, rel))
        self.data_start = (off + self.alignment - 1) // self.alignment * self.alignment
        for name, ne, tt, rel in infos:
            if name in self.tensors:
                raise GGUFError(f"{self.path}: duplicate tensor {name}")
            if tt in GGML_TYPES:
                rows = 1
                for d in ne[1:]:
                    rows *= d
                nbytes = rows * row_bytes(tt, ne[0]) if ne else 0
            else:
                nbytes = -1  # unknown type: listed, but cannot be sized or loaded
            absoff = self.data_start + rel
            if rel % self.alignment:
                raise GGUFError(f"{self.path}: tensor {name} offset {rel} not {self.alignment}-byte aligned")
            if nbytes >= 0:
                self._need(absoff, nbytes)
            self.tensors[name] = TensorInfo(name, tt, ne, absoff, nbytes)

    # -- access
    def data(self, name: str) -> np.ndarray:
        t = self.tensors[name]
        if t.nbytes < 0:
            raise GGUFError(f"tensor {name}: unsupported type {t.type_name}")
        return np.frombuffer(self._buf, dtype=np.uint8, count=t.nbytes, offset=t.offset)

    def get(self, key: str, default=None):
        return self.meta.get(key, default)

    def close(self):
        try:
            self._buf.release()
            self._mm.close()
        except (BufferError, ValueError):
            pass  # numpy views still alive; the mapping goes when they do
        self._f.close()


# ----------------------------------------------------------------------------- writer
class GGUFWriter:
    """Minimal GGUF v3 writer (used for the synthetic models; layout identical to gguf.GGUFWriter's)."""

    def __init__(self, alignment: int = DEFAULT_ALIGNMENT):
        self.alignment = alignment
        self._kv: list[bytes] = []
        self._tensors: list[tuple] = []   # (name, ne, type, bytes-or-callable, size)

    @staticmethod
    def _s(s: str) -> bytes:
        b = s.encode("utf-8")
        return struct.pack("<Q", len(b)) + b

    def add(self, key: str, vtype: int, value):
        out = self._s(key) + struct.pack("<I", vtype)
        if vtype == T_STR:
            out += self._s(value)
        elif vtype in _SCALAR:
            out += struct.pack(_SCALAR[vtype], value)
        else:
            raise GGUFError("use add_array for arrays")
        self._kv.append(out)

    def add_array(self, key: str, etype: int, values):
        out = self._s(key) + struct.pack("<I", T_ARR) + struct.pack("<IQ", etype, len(values))
        if etype == T_STR:
            out += b"".join(self._s(v) for v in values)
        else:
            out += np.asarray(values, dtype=np.dtype(_NP[etype]).newbyteorder("<")).tobytes()
        self._kv.append(out)

    def add_tensor(self, name: str, ne: tuple, ggml_type: int, raw):
        """ne is in ggml order (innermost first); raw = the tensor's bytes (any dtype, C-contiguous), or a
        zero-argument callable producing them at write time (keeps a 40 GB model out of host memory)."""
        rows = 1
        for d in ne[1:]:
            rows *= d
        size = rows * row_bytes(ggml_type, ne[0])
        if not callable(raw):
            raw = np.ascontiguousarray(raw).reshape(-1).view(np.uint8)
            if raw.size != size:
                raise GGUFError(f"{name}: {raw.size} bytes given, {size} expected")
        self._tensors.append((name, tuple(ne), ggml_type, raw, size))

    def write(self, path: str):
        a = self.alignment
        head = struct.pack("<IIQQ", GGUF_MAGIC, 3, len(self._tensors), len(self._kv)) + b"".join(self._kv)
        infos, rel = [], 0
        for name, ne, tt, _, size in self._tensors:
            infos.append(self._s(name) + struct.pack("<I", len(ne)) + b"".join(struct.pack("<Q", d) for d in ne)
                         + struct.pack("<IQ", tt, rel))
            rel = (rel + size + a - 1) // a * a
        head += b"".join(infos)
        with open(path, "wb") as f:
            f.write(head)
            f.write(b"\0" * ((-len(head)) % a))
            # lazily produced tensors are generated a few ahead on worker threads, written in order
            from concurrent.futures import ThreadPoolExecutor
            ahead = max(1, min(6, (os.cpu_count() or 2) - 1))
            with ThreadPoolExecutor(ahead) as pool:
                pending = {}
                for i, (name, _, _, raw, size) in enumerate(self._tensors):
                    for j in range(i, min(i + ahead, len(self._tensors))):
                        if j not in pending and callable(self._tensors[j][3]):
                            pending[j] = pool.submit(self._tensors[j][3])
                    if callable(raw):
                        raw = np.ascontiguousarray(pending.pop(i).result()).reshape(-1).view(np.uint8)
                        if raw.size != size:
                            raise GGUFError(f"{name}: {raw.size} bytes produced, {size} expected")
                    f.write(raw.data)
                    f.write(b"\0" * ((-size) % a))
