"""ctypes binding of oracle/libgiddy_oracle.so -- TEST INFRASTRUCTURE ONLY.

The scalar host decoders of oracle/giddy_oracle.cpp over numpy arrays.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
nothing under libgiddy_b200/ does.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libgiddy_oracle.so")
_lib = None

MODEL_CONSTANT, MODEL_LINEAR = 0, 1
_MODEL = {"constant": MODEL_CONSTANT, "linear": MODEL_LINEAR}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError("%s has not been built (run `make oracle`)" % _PATH)
        _lib = ctypes.CDLL(_PATH)
    return _lib


def use_native_build():
    """For the timed CPU baseline: rebuild the oracle -march=native ON THIS MACHINE (the committed recipe builds it
    -march=x86-64-v3 so that the binary runs on any box it travels to) and switch to it.  Returns the -march used;
    without a compiler here the portable build stays in use."""
    global _lib
    import subprocess
    native = os.path.join(_HERE, "libgiddy_oracle_native.so")
    try:
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", native,
                        os.path.join(_HERE, "giddy_oracle.cpp")], check=True, capture_output=True, timeout=300)
        _lib = ctypes.CDLL(native)
        return "native"
    except Exception:
        return "x86-64-v3"


def num_threads():
    return lib().giddy_oracle_num_threads()


def set_num_threads(n):
    lib().giddy_oracle_set_num_threads(int(n))


def _p(a):
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def _u64(x):
    return ctypes.c_uint64(int(x))


def _check(status, what):
    if status != 0:
        raise ValueError("oracle %s: unsupported or invalid parameters (status %d)" % (what, status))


_UINT = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
_INT = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}


def frame_of_reference(compressed, coefficients, n, segment_length, uncompressed_size, model="constant", model_index_size=4):
    out = np.empty(n, dtype=_INT[uncompressed_size])
    _check(lib().giddy_oracle_frame_of_reference(_p(out), _p(compressed), _p(coefficients), _u64(n), _u64(segment_length),
           uncompressed_size, compressed.dtype.itemsize, _MODEL[model], model_index_size), "frame_of_reference")
    return out


def model(coefficients, n, uncompressed_size, model="linear", index_size=4):
    out = np.empty(n, dtype=_INT[uncompressed_size])
    coefficients = np.ascontiguousarray(coefficients, dtype=_INT[uncompressed_size])
    _check(lib().giddy_oracle_model(_p(out), _p(coefficients), _u64(n), uncompressed_size, _MODEL[model], index_size), "model")
    return out


def run_length_encoding(run_data, run_lengths, anchors, intra_offsets, segment_length, n, out=None):
    num_segments = (n + segment_length - 1) // segment_length
    if out is None:
        out = np.empty(n, dtype=run_data.dtype)
    _check(lib().giddy_oracle_run_length_encoding(_p(out), _p(run_data), _p(run_lengths), _p(anchors), _p(intra_offsets),
           _u64(segment_length), _u64(num_segments), _u64(len(run_data)), _u64(n),
           anchors.dtype.itemsize, run_data.dtype.itemsize, run_lengths.dtype.itemsize), "run_length_encoding")
    return out


def run_position_encoding(run_data, run_starts, anchors, segment_length, n, relative):
    num_segments = (n + segment_length - 1) // segment_length
    out = np.empty(n, dtype=run_data.dtype)
    _check(lib().giddy_oracle_run_position_encoding(_p(out), _p(run_data), _p(run_starts), _p(anchors),
           _u64(segment_length), _u64(num_segments), _u64(len(run_data)), _u64(n),
           anchors.dtype.itemsize, run_data.dtype.itemsize, run_starts.dtype.itemsize, 1 if relative else 0), "run_position_encoding")
    return out


def delta(compressed, anchors, n, segment_length, out=None):
    if out is None:
        out = np.empty(n, dtype=anchors.dtype)
    _check(lib().giddy_oracle_delta(_p(out), _p(compressed), _p(anchors), _u64(n), _u64(segment_length),
           anchors.dtype.itemsize, compressed.dtype.itemsize), "delta")
    return out


def dictionary(compressed, entries, n):
    out = np.empty(n, dtype=entries.dtype)
    _check(lib().giddy_oracle_dictionary(_p(out), _p(compressed), _p(entries), _u64(n),
           entries.dtype.itemsize, compressed.dtype.itemsize), "dictionary")
    return out


def discard_zero_bytes_fixed(compressed, n, uncompressed_size):
    out = np.empty(n, dtype=_UINT[uncompressed_size])
    _check(lib().giddy_oracle_discard_zero_bytes_fixed(_p(out), _p(compressed), _u64(n), uncompressed_size, compressed.dtype.itemsize),
           "discard_zero_bytes_fixed")
    return out


def discard_zero_bytes_variable(data_bytes, packed_sizes, anchors, segment_length, n, min_size, bits, uncompressed_size, out=None):
    if out is None:
        out = np.empty(n, dtype=_UINT[uncompressed_size])
    _check(lib().giddy_oracle_discard_zero_bytes_variable(_p(out), _p(data_bytes), _p(packed_sizes), _p(anchors),
           _u64(segment_length), _u64(n), int(min_size), int(bits),
           anchors.dtype.itemsize, uncompressed_size, packed_sizes.dtype.itemsize), "discard_zero_bytes_variable")
    return out


def incidence_bitmaps(bitmaps, entries, bitmap_length, num_bitmaps, uncompressed_size, alignment_in_ints=2, use_dictionary=True):
    out = np.empty(bitmap_length * 32, dtype=_UINT[uncompressed_size])
    _check(lib().giddy_oracle_incidence_bitmaps(_p(out), _p(bitmaps), _p(entries) if entries is not None else _p(None),
           _u64(bitmap_length), int(num_bitmaps), uncompressed_size, alignment_in_ints, 1 if use_dictionary else 0), "incidence_bitmaps")
    return out


def scatter(target, data, indices):
    """In place on `target`."""
    _check(lib().giddy_oracle_scatter(_p(target), _p(data), _p(indices), _u64(len(data)),
           indices.dtype.itemsize, target.dtype.itemsize), "scatter")
    return target


def compressed_indices_scatter(target, data, run_lengths, offset_sizes, baselines, offsets_start, offset_bytes, anchors, anchoring_period):
    """In place on `target`."""
    _check(lib().giddy_oracle_compressed_indices_scatter(_p(target), _p(data), _p(run_lengths), _p(offset_sizes), _p(baselines),
           _p(offsets_start), _p(offset_bytes), _p(anchors), _u64(anchoring_period), _u64(len(run_lengths)), _u64(len(data)),
           baselines.dtype.itemsize, target.dtype.itemsize, anchors.dtype.itemsize), "compressed_indices_scatter")
    return target
