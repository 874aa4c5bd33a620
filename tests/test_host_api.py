"""CPU tests of the boundary: the shared library loads and exports exactly what include/giddy_b200.h
declares, the factory holds the reference's 163 keys, the launch-configuration resolver behaves as
specified, and errors map to the reference's exception types.  No kernel is launched here."""
import os
import re
import subprocess

import pytest

import libgiddy_b200 as g
from libgiddy_b200 import schemes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROPS = g.b200_device_properties()
ATTRS = g.nominal_function_attributes()


def resolve(library, key, arguments, **kwargs):
    kernel = library.produce_subclass(key)
    return kernel.resolve_launch_configuration(arguments, device_properties=PROPS, kernel_function_attributes=ATTRS, **kwargs)


def test_library_exports_every_declared_symbol(library):
    header = open(os.path.join(ROOT, "include", "giddy_b200.h")).read()
    declared = set(re.findall(r"GIDDY_API[^;(]*?\b(giddy_\w+)\s*\(", header))
    # 42 of the registry, kernel and fused-decoder groups + 12 of the host-resident scan group + 5 shard planning + 4 host memory + 7 device groups
    # + 16 compressors (8 schemes) + 6 utility kernels
    assert len(declared) == 92, len(declared)
    exported = subprocess.run(["nm", "-D", "--defined-only", library.path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (giddy_\w+)", exported))
    assert declared <= exported, declared - exported
    assert library.version().startswith("giddy_b200")


def expected_keys():
    """SURVEY.md Appendix C, expanded (reference src/kernel_wrappers/**/*.cu static_blocks)."""
    model = "cuda::functors::unary::parametric_model::"
    keys = []
    uc = [(2, 1), (4, 1), (4, 2), (8, 1), (8, 2), (8, 4)]
    for i in (4, 8):
        keys += ["decompression::delta::kernel_t<%du, %du, %du>" % (i, u, c) for u, c in uc]
        keys += ["decompression::discard_zero_bytes::fixed_width::kernel_t<%du, %du, %du, (util::endianness_t)1, (unsigned short)32>" % (i, u, c) for u, c in uc]
        keys += ["decompression::dictionary::kernel_t<%du, %du, %du>" % (i, u, d)
                 for u, d in [(1, 1), (2, 1), (2, 2), (4, 1), (4, 2), (4, 4), (8, 1), (8, 2), (8, 4), (8, 8)]]
    keys += ["decompression::discard_zero_bytes::variable_width::kernel_t<%du, %du, %du>" % t
             for t in [(4, 2, 4), (4, 4, 4), (4, 8, 4), (8, 2, 4), (8, 4, 4), (8, 8, 4), (4, 2, 8), (4, 4, 8), (4, 8, 8)]]
    for i, unc, comp, m, mi in [(4, "short", "signed char", "linear", 4), (4, "int", "signed char", "linear", 4), (4, "int", "short", "linear", 4),
                                (8, "int", "short", "linear", 4), (8, "int", "short", "linear", 8), (4, "short", "signed char", "constant", 4),
                                (4, "int", "signed char", "constant", 4), (4, "int", "short", "constant", 4)]:
        keys.append("decompression::frame_of_reference::kernel_t<%du, %s, %s, %s%s<%du, int> >" % (i, unc, comp, model, m, mi))
    keys += ["decompression::incidence_bitmaps::kernel_t<4u, %du, 2u, %s>" % t for t in [(1, "true"), (2, "true"), (4, "true"), (8, "true"), (1, "false")]]
    for i, t, m in [(4, "short", "linear"), (4, "int", "linear"), (8, "int", "linear"), (4, "short", "constant"), (4, "int", "constant")]:
        keys.append("decompression::model::kernel_t<%du, %s, %s%s<%du, %s> >" % (i, t, model, m, i, t))
    keys += ["decompression::run_length_encoding::kernel_t<4u, %du, %du, 4u>" % (u, rl) for rl in (1, 2, 4) for u in (1, 2, 4, 8)]
    keys += ["decompression::run_length_encoding::kernel_t<8u, %du, %du, 8u>" % t for t in [(4, 1), (4, 2), (4, 4), (4, 8), (1, 8), (8, 8)]]
    keys += ["decompression::run_position_encoding::kernel_t<4u, %du, 2u, true, 4u>" % u for u in (1, 2, 4, 8)]
    keys += ["decompression::run_position_encoding::kernel_t<8u, 4u, 2u, true, 8u>"]
    keys += ["decompression::run_position_encoding::kernel_t<4u, %du, 4u, false, 4u>" % u for u in (1, 2, 4, 8)]
    keys += ["decompression::run_position_encoding::kernel_t<8u, 4u, 8u, false, 8u>"]
    for oi in (4, 8):
        for e in (1, 2, 4, 8):
            for ii in (1, 2, 4, 8):
                keys.append("scatter::kernel_t<%du, %du, %du, (unsigned short)16>" % (oi, e, ii))
                keys.append("scatter::compressed_indices::kernel_t<%du, %du, %du, %du>" % (oi, e, ii, ii))
    return keys


def test_factory_holds_the_reference_key_set(library):
    expected = expected_keys()
    assert len(expected) == len(set(expected)) == 163
    registered = library.registered_keys()
    # every reference key, plus nothing but this library's documented extensions
    assert set(expected) <= set(registered)
    # (the compressors and the reference's unwrapped utility kernels have key families of their own, checked in
    # tests/test_compressors.py and tests/test_utility_kernels.py; the reference's namespaces gain composite decoders only)
    def other_family(key):
        return key.startswith(("compression::", "gather::", "decompression::model::intervals::", "reduction::scan::"))
    extensions = sorted(k for k in set(registered) - set(expected) if not other_family(k))
    import libgiddy_b200.schemes as S
    documented = ["decompression::composite::delta_of_discard_zero_bytes_variable_width::kernel_t<%du, %du, %du, %du>" % t
                  for t in [(4, 8, 4, 4), (4, 8, 4, 8), (4, 8, 2, 4), (4, 4, 2, 4), (8, 8, 4, 4)]]
    documented += [S.patched_dictionary_key(*t) for t in [(4, 1, 1), (4, 2, 1), (4, 4, 1), (4, 8, 1), (4, 2, 2), (4, 4, 2), (4, 8, 2),
                                                          (4, 8, 4), (8, 8, 1), (8, 4, 2)]]
    documented += [S.patched_frame_of_reference_key(*t) for t in [(4, 2, 1, "linear", 4, 4), (4, 4, 1, "linear", 4, 4), (4, 4, 2, "linear", 4, 4),
                                                                  (8, 4, 2, "linear", 8, 8), (4, 2, 1, "constant", 4, 4),
                                                                  (4, 4, 1, "constant", 4, 4), (4, 4, 2, "constant", 4, 4)]]
    assert extensions == sorted(documented)
    # the key the reference's own example uses (examples/src/modes_of_use/factory_provided_type_erased_wrapper.cu:15-16)
    example = "decompression::frame_of_reference::kernel_t<4u, int, short, cuda::functors::unary::parametric_model::constant<4u, int> >"
    assert library.can_produce(example)
    assert library.produce_subclass(example).key == example


def test_key_builders_cover_the_factory(library):
    keys = set(library.registered_keys())
    assert S.frame_of_reference_key(4, 4, 2) in keys and S.model_key(8, 4, "linear") in keys
    assert S.run_length_encoding_key(4, 4, 2) in keys and S.run_position_encoding_key(8, 4, 8, False) in keys
    assert S.delta_key(4, 8, 4) in keys and S.dictionary_key(4, 8, 1) in keys
    assert S.discard_zero_bytes_fixed_key(8, 8, 4) in keys and S.discard_zero_bytes_variable_key(4, 4, 4) in keys
    assert S.incidence_bitmaps_key(4, 4) in keys and S.scatter_key(4, 8, 4) in keys and S.compressed_indices_scatter_key(8, 1, 2) in keys


def test_unknown_key_is_invalid_argument(library):
    with pytest.raises(g.InvalidArgument):
        library.produce_subclass("decompression::delta::kernel_t<4u, 3u, 1u>")
    assert not library.can_produce("nonsense")


def test_missing_and_mistyped_arguments(library):
    key = S.delta_key(4, 8, 4)
    with pytest.raises(g.OutOfRange):
        resolve(library, key, {"length": g.size_t(1000)})
    with pytest.raises(g.BadAnyCast):  # resolve-map scalars are size_t, not uint32_t (reference delta.cu:55-56)
        resolve(library, key, {"length": g.u32(1000), "segment_length": g.size_t(10)})
    kernel = library.produce_subclass(key)
    config = g.make_launch_config(1, 32)
    with pytest.raises(g.OutOfRange):
        kernel.enqueue_launch(0, config, {"decompressed": g.ptr(0, "u64")})
    with pytest.raises(g.BadAnyCast):  # launch-map length is uint_t<IndexSize> (reference delta.cu:76-77); also const-ness matters
        kernel.enqueue_launch(0, config, {"decompressed": g.ptr(0, "u64"), "compressed_input": g.ptr(0, "u32"),
                                          "anchor_values": g.cptr(0, "u64"), "length": g.u32(1), "segment_length": g.u32(1)})


def test_scheme_parameter_validation(library):
    with pytest.raises(g.InvalidArgument):  # reference frame_of_reference.cuh:99-101
        resolve(library, S.frame_of_reference_key(4, 4, 2), {"length": g.size_t(100), "segment_length": g.size_t(0)})
    with pytest.raises(g.InvalidArgument):  # reference frame_of_reference.cuh:102-104
        resolve(library, S.frame_of_reference_key(4, 4, 2), {"length": g.size_t(100), "segment_length": g.size_t(6)})
    with pytest.raises(g.InvalidArgument):  # reference compressed_indices_scatter.cuh:288-290
        resolve(library, S.compressed_indices_scatter_key(4, 4, 4), {"input_data_length": g.size_t(0), "anchoring_period": g.size_t(64)})


def test_resolved_configurations_for_the_baseline_shapes(library):
    """The B200 launch policy at BASELINE.json's sizes ('Launch-configuration policy')."""
    c = resolve(library, S.frame_of_reference_key(4, 4, 2), {"length": g.size_t(1 << 26), "segment_length": g.size_t(4096)})
    assert (c.grid, c.block, c.dynamic_shared_memory_size) == ((1 << 26) // (4 * 4 * 256), 256, 0)
    c = resolve(library, S.run_length_encoding_key(4, 4, 2), {"length": g.size_t(1 << 30), "segment_length": g.size_t(1024)})
    assert (c.grid, c.block, c.dynamic_shared_memory_size) == ((1 << 20) // 4, 128, 4 * (256 + 1056 * 4))
    c = resolve(library, S.delta_key(4, 8, 4), {"length": g.size_t(1 << 30), "segment_length": g.size_t(1024)})
    assert (c.grid, c.block) == ((1 << 20) // 8, 256)
    c = resolve(library, S.dictionary_key(4, 8, 1), {"num_dictionary_entries": g.size_t(256), "data_length": g.size_t(1 << 29)})
    assert c.block == 256 and c.dynamic_shared_memory_size == 2048
    c = resolve(library, S.dictionary_key(4, 8, 1), {"num_dictionary_entries": g.size_t(256), "data_length": g.size_t(1 << 29),
                                                      "cache_input_data_in_shared_mem": g.boolean(False)})
    assert c.dynamic_shared_memory_size == 0


def test_limits_are_honoured(library):
    key = S.discard_zero_bytes_fixed_key(4, 4, 2)
    args = {"length": g.size_t(1 << 20)}
    c = resolve(library, key, args, limits={"block_size": 64})
    assert c.block == 64 and c.grid == (1 << 20) // (16 * 64)
    with pytest.raises(g.InvalidArgument):  # not an auto-serialising kernel: cannot shrink the grid (reference resolve_launch_configuration.cpp:320-327)
        resolve(library, key, args, limits={"grid_size": 10})
    c = resolve(library, S.run_length_encoding_key(4, 8, 2), {"length": g.size_t(1 << 20), "segment_length": g.size_t(1024)},
                limits={"dynamic_shared_memory": 20000})
    assert c.block == 64 and c.dynamic_shared_memory_size == 2 * (256 + 1056 * 8)
    with pytest.raises(g.InvalidArgument):  # no room for even one warp
        resolve(library, S.run_length_encoding_key(4, 8, 2), {"length": g.size_t(1 << 20), "segment_length": g.size_t(1024)},
                limits={"dynamic_shared_memory": 1000})


def test_tiny_inputs_resolve_to_one_small_block(library):
    c = resolve(library, S.model_key(4, 4, "linear"), {"length": g.size_t(5)})
    assert (c.grid, c.block) == (1, 32)
    c = resolve(library, S.delta_key(4, 8, 4), {"length": g.size_t(5), "segment_length": g.size_t(1024)})
    assert (c.grid, c.block) == (1, 32)


def test_dictionary_serialization_factor_argument(library):
    key = S.dictionary_key(4, 4, 2)
    base = {"num_dictionary_entries": g.size_t(60000), "data_length": g.size_t(1 << 24), "cache_input_data_in_shared_mem": g.boolean(False)}
    c1 = resolve(library, key, base)
    c2 = resolve(library, key, dict(base, serialization_factor=g.serialization_factor(64)))
    assert c1.grid == (1 << 24) // (16 * 256) and c2.grid == (1 << 24) // (64 * 256)


def test_key_set_equals_the_rebuilt_reference_factory(library, reference_library):
    """The reference's own registration code (static_blocks + abi::__cxa_demangle), rebuilt from
    /root/reference by oracle/build_ref.sh, yields exactly the product's key set (less the product's
    documented extension families)."""
    extension_families = ("decompression::composite::", "compression::", "gather::", "decompression::model::intervals::", "reduction::scan::")
    ours = [k for k in library.registered_keys() if not k.startswith(extension_families)]
    assert sorted(reference_library.registered_keys()) == sorted(ours)


def test_policies_of_the_persistent_kernels(library):
    """Large dictionaries, long Delta segments and the patched decoders run persistent grids sized from the device."""
    sms = PROPS.multiprocessor_count
    # a 240 KB dictionary and a column 1000x its size: one 1024-thread CTA per multiprocessor holding as much as fits
    c = resolve(library, S.dictionary_key(4, 4, 2), {"num_dictionary_entries": g.size_t(60000), "data_length": g.size_t(1 << 27)})
    assert (c.grid, c.block) == (sms, 1024) and 200 * 1024 < c.dynamic_shared_memory_size <= 227 * 1024 and c.dynamic_shared_memory_size % 16 == 0
    # the same dictionary for a short column: gathered through L1/L2
    c = resolve(library, S.dictionary_key(4, 4, 2), {"num_dictionary_entries": g.size_t(60000), "data_length": g.size_t(1 << 22)})
    assert c.block == 256 and c.dynamic_shared_memory_size == 0
    # 160 KB: all of it in shared memory
    c = resolve(library, S.dictionary_key(4, 4, 2), {"num_dictionary_entries": g.size_t(40000), "data_length": g.size_t(1 << 27)})
    assert (c.grid, c.block, c.dynamic_shared_memory_size) == (sms, 1024, 160000)
    # Delta over one long segment: tiles + decoupled look-back, 2 CTAs of 256 threads per multiprocessor
    c = resolve(library, S.delta_key(4, 8, 4), {"length": g.size_t(1 << 26), "segment_length": g.size_t(1 << 26)})
    assert (c.grid, c.block, c.dynamic_shared_memory_size) == (2 * sms, 256, 0)
    # ... but anchored segments stay with a warp per segment
    c = resolve(library, S.delta_key(4, 8, 4), {"length": g.size_t(1 << 26), "segment_length": g.size_t(4096)})
    assert (c.grid, c.block) == ((1 << 14) // 8, 256)
    # patched decoders: 4 CTAs of 256 threads per multiprocessor, warps striding over the tiles
    c = resolve(library, S.patched_dictionary_key(4, 8, 1), {"num_dictionary_entries": g.size_t(256), "data_length": g.size_t(1 << 29)})
    assert (c.grid, c.block, c.dynamic_shared_memory_size) == (4 * sms, 256, 2048)
    c = resolve(library, S.patched_dictionary_key(4, 8, 1), {"num_dictionary_entries": g.size_t(256), "data_length": g.size_t(5000)})
    assert c.grid == 3 and c.block == 256  # 5000 elements: three CTA tiles of 256 lanes x 2 pieces x 4 elements (32-byte pieces)
