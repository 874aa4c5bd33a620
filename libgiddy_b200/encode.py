"""Host-side (numpy) compressors for the Giddy formats.

The reference ships decompressors only (README.md:128-130, its Issue #3), so everything that feeds a
decompressor -- tests, benchmarks, examples -- needs compressors with anchor generation.  These
produce exactly the layouts of SURVEY.md Appendix A, vectorised with numpy.  They are host utilities,
not part of the GPU decode path.
"""
import numpy as np

UINT = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
INT = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}


def num_segments(n, segment_length):
    return (n + segment_length - 1) // segment_length


def _runs_of(values):
    """Start positions and lengths of the maximal runs of equal adjacent values."""
    n = len(values)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    starts = np.concatenate(([0], np.flatnonzero(values[1:] != values[:-1]) + 1)).astype(np.int64)
    lengths = np.diff(np.concatenate((starts, [n])))
    return starts, lengths


def _split_long_runs(starts, lengths, max_length):
    """Break runs longer than max_length into consecutive pieces."""
    if len(lengths) == 0 or lengths.max() <= max_length:
        return starts, lengths
    pieces = (lengths + max_length - 1) // max_length
    owner = np.repeat(np.arange(len(lengths)), pieces)
    piece_index = np.arange(len(owner)) - np.repeat(np.cumsum(pieces) - pieces, pieces)
    new_starts = starts[owner] + piece_index * max_length
    new_lengths = np.minimum(max_length, lengths[owner] - piece_index * max_length)
    return new_starts, new_lengths


def run_length_encode(values, segment_length, run_length_size=2, index_size=4, max_run_length=None):
    """Appendix A.3.  Returns dict(run_data, run_lengths, position_anchors, intra_run_anchor_offsets)."""
    values = np.ascontiguousarray(values)
    limit = (1 << (8 * run_length_size)) - 1
    max_run_length = min(max_run_length or limit, limit)
    starts, lengths = _split_long_runs(*_runs_of(values), max_run_length)
    segment_starts = np.arange(num_segments(len(values), segment_length), dtype=np.int64) * segment_length
    anchors = np.searchsorted(starts, segment_starts, side="right") - 1
    return {
        "run_data": values[starts].copy(),
        "run_lengths": lengths.astype(UINT[run_length_size]),
        "position_anchors": anchors.astype(UINT[index_size]),
        "intra_run_anchor_offsets": (segment_starts - starts[anchors]).astype(UINT[run_length_size]),
    }


def run_position_encode(values, segment_length, position_offset_size=4, relative=False, index_size=4, max_run_length=None):
    """Appendix A.4.  Relative positions are relative to the segment in which the run begins."""
    values = np.ascontiguousarray(values)
    starts, lengths = _runs_of(values)
    if max_run_length:
        starts, lengths = _split_long_runs(starts, lengths, max_run_length)
    segment_starts = np.arange(num_segments(len(values), segment_length), dtype=np.int64) * segment_length
    anchors = np.searchsorted(starts, segment_starts, side="right") - 1
    stored = starts % segment_length if relative else starts
    if len(stored) and stored.max() >= (1 << (8 * position_offset_size)):
        raise ValueError("run start positions do not fit in %d bytes" % position_offset_size)
    return {
        "run_data": values[starts].copy(),
        "run_start_positions": stored.astype(UINT[position_offset_size]),
        "position_anchors": anchors.astype(UINT[index_size]),
    }


def delta_encode(values, segment_length, compressed_size, base=0):
    """Appendix A.5.  values: unsigned; every difference (mod 2^(8U)) must fit in compressed_size bytes.
    anchor[s] is the value preceding segment s (`base` for the first)."""
    values = np.ascontiguousarray(values)
    u = values.dtype
    previous = np.concatenate((np.array([base], dtype=u), values[:-1]))
    deltas = (values - previous).astype(u)  # wraps
    if len(deltas) and int(deltas.max()) >= (1 << (8 * compressed_size)):
        raise ValueError("a delta does not fit in %d bytes" % compressed_size)
    return {
        "compressed_input": deltas.astype(UINT[compressed_size]),
        "anchor_values": previous[::segment_length].copy(),
    }


def significant_bytes(values):
    """Number of bytes up to and including the most significant non-zero one (0 for the value 0)."""
    v = values.astype(np.uint64)
    sizes = np.zeros(len(v), dtype=np.int64)
    for k in range(values.dtype.itemsize):
        sizes[(v >> np.uint64(8 * k)) != 0] = k + 1
    return sizes


def discard_zero_bytes_variable_encode(values, segment_length, container_size=4, bits_per_element_size=2,
                                       min_represented_element_size=1, index_size=4):
    """Appendix A.8.  Returns dict(compressed_data, packed_element_sizes, position_anchors) + the sizes."""
    values = np.ascontiguousarray(values)
    n, u = len(values), values.dtype.itemsize
    sizes = np.maximum(significant_bytes(values), min_represented_element_size)
    if len(sizes) and sizes.max() - min_represented_element_size >= (1 << bits_per_element_size):
        raise ValueError("an element size is not representable")
    byte_matrix = values.view(np.uint8).reshape(n, u)
    keep = np.arange(u)[None, :] < sizes[:, None]
    compressed = byte_matrix[keep]  # row-major: element order, then byte order
    spc = (8 * container_size) // bits_per_element_size
    fields = np.zeros(((n + spc - 1) // spc) * spc, dtype=np.uint64)
    fields[:n] = sizes - min_represented_element_size
    shifts = (np.arange(spc, dtype=np.uint64) * np.uint64(bits_per_element_size))[None, :]
    containers = np.bitwise_or.reduce(fields.reshape(-1, spc) << shifts, axis=1).astype(UINT[container_size])
    offsets = np.concatenate(([0], np.cumsum(sizes)))
    anchors = offsets[np.arange(num_segments(n, segment_length)) * segment_length]
    if len(anchors) and anchors.max() >= (1 << (8 * index_size)):
        raise ValueError("byte offsets do not fit the index type")
    return {
        "compressed_data": np.ascontiguousarray(compressed),
        "packed_element_sizes": containers,
        "position_anchors": anchors.astype(UINT[index_size]),
        "element_sizes": sizes,
    }


def incidence_bitmaps_encode(indices, num_bitmaps, alignment_in_ints=2):
    """Appendix A.9.  indices: the bitmap (value) index of every element.  Returns (bitmaps, bitmap_length)."""
    n = len(indices)
    bitmap_length = (n + 31) // 32
    aligned = (bitmap_length + alignment_in_ints - 1) // alignment_in_ints * alignment_in_ints
    bitmaps = np.zeros(num_bitmaps * aligned, dtype=np.uint32)
    padded = np.full(bitmap_length * 32, -1, dtype=np.int64)
    padded[:n] = indices
    for b in range(num_bitmaps):
        bits = np.packbits(padded == b, bitorder="little").view(np.uint32)
        bitmaps[b * aligned: b * aligned + bitmap_length] = bits
    return bitmaps, bitmap_length


def frame_of_reference_encode(values, segment_length, compressed_size, model="constant"):
    """Appendix A.1, constant model: the reference value of a segment is its minimum plus the
    compressed type's bias, so offsets span the signed compressed range; linear model: the line
    through the segment's first and last elements, offsets being the residuals.
    Raises if an offset does not fit.  values: signed integers."""
    values = np.ascontiguousarray(values)
    n = len(values)
    S = num_segments(n, segment_length)
    padded = np.empty(S * segment_length, dtype=np.int64)
    padded[:n] = values
    padded[n:] = values[-1] if n else 0
    seg = padded.reshape(S, segment_length)
    lo, hi = -(1 << (8 * compressed_size - 1)), (1 << (8 * compressed_size - 1)) - 1
    if model == "constant":
        c0 = seg.min(axis=1) - lo
        coefficients = c0[:, None].astype(np.int32)
        predicted = np.repeat(c0, segment_length)
    else:
        first, last = seg[:, 0], seg[:, -1]
        slope = np.round((last - first) / max(segment_length - 1, 1)).astype(np.int64)
        x = np.arange(segment_length, dtype=np.int64)
        residual = seg - (first[:, None] + slope[:, None] * x[None, :])
        c0 = first + (residual.min(axis=1) - lo)
        coefficients = np.stack((c0, slope), axis=1).astype(np.int32)
        predicted = (c0[:, None] + slope[:, None] * x[None, :]).reshape(-1)
    offsets = padded - predicted
    if offsets[:n].min(initial=0) < lo or offsets[:n].max(initial=0) > hi:
        raise ValueError("an offset does not fit the compressed type")
    return {"compressed_input": offsets[:n].astype(INT[compressed_size]), "segment_model_coefficients": np.ascontiguousarray(coefficients)}


def dictionary_encode(values, dictionary_index_size=1):
    entries, indices = np.unique(values, return_inverse=True)
    if len(entries) > (1 << (8 * dictionary_index_size)):
        raise ValueError("too many distinct values for the dictionary index type")
    return {"dictionary_entries": entries, "compressed_input": indices.astype(UINT[dictionary_index_size])}


def compressed_indices_encode(positions, anchoring_period, output_index_size=4, input_index_size=4, max_run_length=64,
                              allowed_widths=(8, 16, 32, 64)):
    """Appendix A.11.  positions: sorted distinct scatter targets.  Patches are cut into runs of at most
    max_run_length which never straddle a multiple of anchoring_period; each run stores its first
    position as the baseline and its offsets in the narrowest allowed width."""
    positions = np.ascontiguousarray(positions).astype(np.int64)
    p = len(positions)
    run_starts = []
    for period_start in range(0, p, anchoring_period):
        period_end = min(period_start + anchoring_period, p)
        run_starts.extend(range(period_start, period_end, max_run_length))
    run_starts = np.array(run_starts, dtype=np.int64)
    period_ends = np.minimum((run_starts // anchoring_period + 1) * anchoring_period, p)
    run_ends = np.minimum(run_starts + max_run_length, period_ends)
    run_lengths = run_ends - run_starts
    baselines = positions[run_starts] if p else np.zeros(0, np.int64)
    max_offsets = np.array([positions[e - 1] - positions[s] for s, e in zip(run_starts, run_ends)], dtype=np.int64)
    widths = np.full(len(run_starts), max(allowed_widths), dtype=np.int64)
    for w in sorted(allowed_widths, reverse=True):
        if w < 64:
            widths[max_offsets < (1 << w)] = w
    chunks, offsets_start, cursor = [], [], 0
    for s, e, w in zip(run_starts, run_ends, widths):
        offsets_start.append(cursor // output_index_size)
        raw = (positions[s:e] - positions[s]).astype(UINT[w // 8]).view(np.uint8)
        padding = (-len(raw)) % output_index_size
        chunks.append(raw)
        if padding:
            chunks.append(np.zeros(padding, dtype=np.uint8))
        cursor += len(raw) + padding
    offset_bytes = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    anchors = np.searchsorted(run_starts, np.arange(0, p, anchoring_period), side="left")
    return {
        "scatter_position_run_lengths": run_lengths.astype(UINT[input_index_size]),
        "scatter_position_run_individual_offset_sizes": widths.astype(np.uint8),
        "scatter_position_run_baseline_values": baselines.astype(UINT[output_index_size]),
        "scatter_position_run_offsets_start_pos": np.array(offsets_start, dtype=np.int64).astype(UINT[input_index_size]),
        "scatter_position_offset_bytes": np.ascontiguousarray(offset_bytes),
        "scatter_position_anchors": anchors.astype(UINT[input_index_size]),
    }
