"""ctypes binding of libgiddy_b200.so's registry API (include/giddy_b200.h, group 1).

This is the Python-side mirror of the reference's "mode iii" interface
(examples/src/modes_of_use/factory_provided_type_erased_wrapper.cu:8-46):

    kernel = produce_subclass("decompression::delta::kernel_t<4u, 8u, 4u>")
    config = kernel.resolve_launch_configuration({"length": size_t(n), "segment_length": size_t(L)})
    kernel.enqueue_launch(stream, config, {"decompressed": ptr(out, "u64"), ...})

Argument values carry their exact static type, as with boost::any in the reference
(src/cuda/kernel_wrapper.cuh:47-49): a mistyped argument raises BadAnyCast, a missing one
OutOfRange, an unknown key InvalidArgument.

There is no fallback: if the shared library has not been built, importing this module fails.
"""
import ctypes
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgiddy_b200.so")


class GiddyError(Exception):
    """Base of the errors raised by libgiddy_b200 (one subclass per giddy_status_t)."""


class InvalidArgument(GiddyError, ValueError):
    """std::invalid_argument in the reference's C++ API"""


class OutOfRange(GiddyError, KeyError):
    """std::out_of_range: a required named argument is missing"""


class BadAnyCast(GiddyError, TypeError):
    """boost::bad_any_cast: a named argument does not have exactly the expected type"""


class LogicError(GiddyError):
    """std::logic_error"""


class CudaError(GiddyError, RuntimeError):
    """cuda::runtime_error"""


class Unsupported(GiddyError, ValueError):
    """the kernel API was asked for a size combination which is not instantiated"""


_STATUS_TO_ERROR = {1: InvalidArgument, 2: OutOfRange, 3: BadAnyCast, 4: LogicError, 5: CudaError, 6: Unsupported, 7: GiddyError}

# giddy_type_t
T_U8, T_U16, T_U32, T_U64, T_I8, T_I16, T_I32, T_I64, T_BOOL = range(1, 10)
T_ARRAY_I16_1, T_ARRAY_I16_2, T_ARRAY_I32_1, T_ARRAY_I32_2 = range(10, 14)
T_PTR, T_CONST = 0x100, 0x200

_TYPE_TAGS = {
    "u8": T_U8, "u16": T_U16, "u32": T_U32, "u64": T_U64, "i8": T_I8, "i16": T_I16, "i32": T_I32, "i64": T_I64, "bool": T_BOOL,
    "array_i16_1": T_ARRAY_I16_1, "array_i16_2": T_ARRAY_I16_2, "array_i32_1": T_ARRAY_I32_1, "array_i32_2": T_ARRAY_I32_2,
}
_SIGNED = {T_I8, T_I16, T_I32, T_I64}
_ARRAYS = {T_ARRAY_I16_1: 1, T_ARRAY_I16_2: 2, T_ARRAY_I32_1: 1, T_ARRAY_I32_2: 2}


class _ArgValue(ctypes.Union):
    _fields_ = [("u", ctypes.c_uint64), ("i", ctypes.c_int64), ("p", ctypes.c_void_p), ("a", ctypes.c_int32 * 2)]


class _Arg(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("type", ctypes.c_int), ("value", _ArgValue)]


class LaunchConfig(ctypes.Structure):
    """giddy_launch_config_t / cuda::launch_configuration_t"""
    _fields_ = [("grid_dimensions", ctypes.c_uint * 3), ("block_dimensions", ctypes.c_uint * 3),
                ("dynamic_shared_memory_size", ctypes.c_uint)]

    @property
    def grid(self):
        return self.grid_dimensions[0]

    @property
    def block(self):
        return self.block_dimensions[0]

    def __repr__(self):
        return "LaunchConfig(grid=%d, block=%d, dynamic_shared_memory=%d)" % (self.grid, self.block, self.dynamic_shared_memory_size)


def make_launch_config(grid, block, dynamic_shared_memory=0):
    c = LaunchConfig()
    c.grid_dimensions[:] = [grid, 1, 1]
    c.block_dimensions[:] = [block, 1, 1]
    c.dynamic_shared_memory_size = dynamic_shared_memory
    return c


class _Limits(ctypes.Structure):
    _fields_ = [("has_block_size", ctypes.c_int), ("block_size", ctypes.c_uint),
                ("has_grid_size", ctypes.c_int), ("grid_size", ctypes.c_uint),
                ("has_dynamic_shared_memory", ctypes.c_int), ("dynamic_shared_memory", ctypes.c_uint)]


class DeviceProperties(ctypes.Structure):
    """giddy_device_properties_t"""
    _fields_ = [("compute_capability_major", ctypes.c_int), ("compute_capability_minor", ctypes.c_int),
                ("multiprocessor_count", ctypes.c_int), ("max_threads_per_block", ctypes.c_int),
                ("max_threads_dim_x", ctypes.c_int), ("max_threads_per_multiprocessor", ctypes.c_int),
                ("max_blocks_per_multiprocessor", ctypes.c_int), ("warp_size", ctypes.c_int),
                ("shared_memory_per_block", ctypes.c_size_t), ("shared_memory_per_block_optin", ctypes.c_size_t),
                ("shared_memory_per_multiprocessor", ctypes.c_size_t), ("l2_cache_size", ctypes.c_size_t),
                ("total_global_memory", ctypes.c_size_t)]


def b200_device_properties():
    """The properties of a B200, for resolving launch configurations where no GPU is present."""
    p = DeviceProperties()
    p.compute_capability_major, p.compute_capability_minor = 10, 0
    p.multiprocessor_count = 148
    p.max_threads_per_block = p.max_threads_dim_x = 1024
    p.max_threads_per_multiprocessor = 2048
    p.max_blocks_per_multiprocessor = 32
    p.warp_size = 32
    p.shared_memory_per_block = 48 * 1024
    p.shared_memory_per_block_optin = 227 * 1024
    p.shared_memory_per_multiprocessor = 228 * 1024
    p.l2_cache_size = 126 * 1024 * 1024
    p.total_global_memory = 180 * 1024 ** 3
    return p


class FunctionAttributes(ctypes.Structure):
    """giddy_function_attributes_t"""
    _fields_ = [("max_threads_per_block", ctypes.c_int), ("num_regs", ctypes.c_int),
                ("shared_size_bytes", ctypes.c_size_t), ("const_size_bytes", ctypes.c_size_t),
                ("local_size_bytes", ctypes.c_size_t), ("max_dynamic_shared_size_bytes", ctypes.c_int),
                ("ptx_version", ctypes.c_int), ("binary_version", ctypes.c_int)]


def nominal_function_attributes(max_threads_per_block=1024):
    a = FunctionAttributes()
    a.max_threads_per_block = max_threads_per_block
    a.max_dynamic_shared_size_bytes = 48 * 1024
    a.ptx_version = a.binary_version = 100
    return a


def load_library(path=_LIB_PATH):
    """Load a shared library exporting the registry API and declare its prototypes."""
    if not os.path.exists(path):
        raise ImportError(
            "%s has not been built (run `make` at the repository root, or __graft_entry__.build()); "
            "libgiddy_b200 has no fallback path" % path)
    lib = ctypes.CDLL(path)
    lib.giddy_last_error.restype = ctypes.c_char_p
    lib.giddy_version.restype = ctypes.c_char_p
    lib.giddy_num_registered_kernels.restype = ctypes.c_size_t
    lib.giddy_registered_kernel_key.restype = ctypes.c_char_p
    lib.giddy_registered_kernel_key.argtypes = [ctypes.c_size_t]
    lib.giddy_can_produce.argtypes = [ctypes.c_char_p]
    lib.giddy_produce_subclass.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.giddy_kernel_destroy.argtypes = [ctypes.c_void_p]
    lib.giddy_kernel_destroy.restype = None
    lib.giddy_kernel_key.argtypes = [ctypes.c_void_p]
    lib.giddy_kernel_key.restype = ctypes.c_char_p
    lib.giddy_kernel_get_device_function.argtypes = [ctypes.c_void_p]
    lib.giddy_kernel_get_device_function.restype = ctypes.c_void_p
    lib.giddy_kernel_resolve_launch_configuration.argtypes = [
        ctypes.c_void_p, ctypes.POINTER(DeviceProperties), ctypes.POINTER(FunctionAttributes),
        ctypes.POINTER(_Arg), ctypes.c_size_t, ctypes.POINTER(_Limits), ctypes.POINTER(LaunchConfig)]
    lib.giddy_kernel_enqueue_launch.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(LaunchConfig), ctypes.POINTER(_Arg), ctypes.c_size_t]
    lib.giddy_device_properties.argtypes = [ctypes.c_int, ctypes.POINTER(DeviceProperties)]
    lib.giddy_function_attributes.argtypes = [ctypes.c_void_p, ctypes.POINTER(FunctionAttributes)]
    if hasattr(lib, "giddy_scan_create"):  # group 4 (host-resident scans): the product only, not the rebuilt reference
        lib.giddy_scan_create.argtypes = [ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p)]
        lib.giddy_scan_destroy.argtypes = [ctypes.c_void_p]
        lib.giddy_scan_destroy.restype = None
        for name in ("giddy_scan_num_streams", "giddy_scan_num_chunks"):
            getattr(lib, name).argtypes = [ctypes.c_void_p]
            getattr(lib, name).restype = ctypes.c_uint
        lib.giddy_scan_num_operations.argtypes = [ctypes.c_void_p]
        lib.giddy_scan_num_operations.restype = ctypes.c_size_t
        lib.giddy_scan_add_copy.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        lib.giddy_scan_add_copy_back.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        lib.giddy_scan_add_launch.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.POINTER(LaunchConfig),
                                              ctypes.POINTER(_Arg), ctypes.c_size_t]
        lib.giddy_scan_add_callback.argtypes = [ctypes.c_void_p, ctypes.c_uint, SCAN_CALLBACK, ctypes.c_void_p, ctypes.c_void_p]
        lib.giddy_scan_add_dependency.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]
        lib.giddy_scan_set_graph_replay.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.giddy_scan_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    if hasattr(lib, "giddy_shard_plan"):  # groups 5-7 (shard planning, host memory, device groups): the product only
        lib.giddy_split_segments.argtypes = [ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        lib.giddy_shard_plan.argtypes = [ctypes.POINTER(ColumnLayout), ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(Shard)]
        lib.giddy_slice_bytes.argtypes = [ctypes.POINTER(Slice)]
        lib.giddy_slice_bytes.restype = ctypes.c_size_t
        lib.giddy_slice_copy.argtypes = [ctypes.POINTER(Slice), ctypes.c_void_p, ctypes.c_void_p]
        lib.giddy_shard_model_coefficients.argtypes = [ctypes.POINTER(ctypes.c_int64), ctypes.c_uint, ctypes.c_uint, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int64)]
        lib.giddy_host_numa_node_of_device.argtypes = [ctypes.c_int]
        lib.giddy_host_bind_thread_to_device.argtypes = [ctypes.c_int]
        lib.giddy_host_alloc.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]
        lib.giddy_host_free.argtypes = [ctypes.c_void_p]
        lib.giddy_device_group_create.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p)]
        lib.giddy_device_group_destroy.argtypes = [ctypes.c_void_p]
        lib.giddy_device_group_destroy.restype = None
        lib.giddy_device_group_size.argtypes = [ctypes.c_void_p]
        lib.giddy_device_group_size.restype = ctypes.c_uint
        lib.giddy_device_group_device.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        lib.giddy_device_group_run.argtypes = [ctypes.c_void_p, DEVICE_WORK, ctypes.c_void_p]
        lib.giddy_device_group_broadcast.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p)]
        lib.giddy_broadcast_dictionary.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    return lib


# giddy_scheme_t
(SCHEME_FRAME_OF_REFERENCE, SCHEME_MODEL, SCHEME_RUN_LENGTH_ENCODING, SCHEME_RUN_POSITION_ENCODING, SCHEME_DELTA, SCHEME_DICTIONARY,
 SCHEME_DISCARD_ZERO_BYTES_FIXED, SCHEME_DISCARD_ZERO_BYTES_VARIABLE, SCHEME_INCIDENCE_BITMAPS, SCHEME_SCATTER) = range(10)


class ColumnLayout(ctypes.Structure):
    """giddy_column_layout_t: what the shard planner needs to know about a whole compressed column"""
    _fields_ = [("scheme", ctypes.c_int), ("length", ctypes.c_uint64), ("segment_length", ctypes.c_uint64),
                ("index_size", ctypes.c_uint), ("uncompressed_size", ctypes.c_uint), ("compressed_size", ctypes.c_uint),
                ("model_dimension", ctypes.c_uint), ("positions_are_relative", ctypes.c_int), ("bits_per_element_size", ctypes.c_uint),
                ("num_runs", ctypes.c_uint64), ("num_bytes", ctypes.c_uint64), ("num_bitmaps", ctypes.c_uint),
                ("bitmap_alignment_in_ints", ctypes.c_uint), ("num_patches", ctypes.c_uint64),
                ("anchors", ctypes.c_void_p), ("run_start_positions", ctypes.c_void_p)]


class Slice(ctypes.Structure):
    """giddy_slice_t"""
    _fields_ = [("array", ctypes.c_char_p), ("first", ctypes.c_uint64), ("count", ctypes.c_uint64), ("element_size", ctypes.c_uint),
                ("rows", ctypes.c_uint64), ("pitch", ctypes.c_uint64), ("rebase", ctypes.c_uint64), ("saturate", ctypes.c_int)]

    @property
    def name(self):
        return self.array.decode()


class Shard(ctypes.Structure):
    """giddy_shard_t"""
    _fields_ = [("first_segment", ctypes.c_uint64), ("num_segments", ctypes.c_uint64), ("first_element", ctypes.c_uint64),
                ("num_elements", ctypes.c_uint64), ("num_runs", ctypes.c_uint64), ("num_bytes", ctypes.c_uint64),
                ("bitmap_length", ctypes.c_uint64), ("num_patches", ctypes.c_uint64), ("num_slices", ctypes.c_uint), ("slices", Slice * 6)]

    def slice(self, name):
        for k in range(self.num_slices):
            if self.slices[k].name == name:
                return self.slices[k]
        raise KeyError(name)

    def names(self):
        return [self.slices[k].name for k in range(self.num_slices)]


# giddy_device_work_t
DEVICE_WORK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, ctypes.c_void_p)


# giddy_scan_callback_t
SCAN_CALLBACK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)


class Value:
    """A named-argument value with its exact static type (the analogue of a boost::any holding a T)."""
    __slots__ = ("tag", "value", "keepalive")

    def __init__(self, tag, value, keepalive=None):
        self.tag, self.value, self.keepalive = tag, value, keepalive

    def __repr__(self):
        return "Value(tag=0x%x, value=%r)" % (self.tag, self.value)


def _scalar(tag):
    def make(x):
        return Value(tag, int(x))
    return make


u8, u16, u32, u64 = _scalar(T_U8), _scalar(T_U16), _scalar(T_U32), _scalar(T_U64)
i8, i16, i32, i64 = _scalar(T_I8), _scalar(T_I16), _scalar(T_I32), _scalar(T_I64)
size_t = u64            # size_t is unsigned long, i.e. uint64_t, on this platform
unsigned = u32
serialization_factor = u16  # cuda::serialization_factor_t is unsigned short


def boolean(x):
    return Value(T_BOOL, 1 if x else 0)


def uint_of_size(size):
    """A scalar constructor for util::uint_t<size>"""
    return {1: u8, 2: u16, 4: u32, 8: u64}[size]


def size_type_by_index_size(size):
    """A scalar constructor for cuda::size_type_by_index_size<size> (32 bits up to 4, else 64)"""
    return u64 if size > 4 else u32


def coefficients(element, values):
    """cuda::array<int16|int32, 1|2> by value (model coefficients)"""
    values = [int(v) for v in values]
    return Value(_TYPE_TAGS["array_%s_%d" % (element, len(values))], values)


def _address_of(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):   # torch.Tensor
        return x.data_ptr()
    if hasattr(x, "ctypes"):     # numpy.ndarray (host memory: only meaningful to a host-side library)
        return x.ctypes.data
    raise TypeError("cannot take the address of a %s" % type(x).__name__)


def ptr(x, element):
    """A T* argument; element names T ('u8', 'i16', 'array_i32_1', ...)"""
    return Value(T_PTR | _TYPE_TAGS[element], _address_of(x), keepalive=x)


def cptr(x, element):
    """A const T* argument"""
    return Value(T_PTR | T_CONST | _TYPE_TAGS[element], _address_of(x), keepalive=x)


def _marshal(arguments):
    items = list(arguments.items())
    array = (_Arg * max(len(items), 1))()
    names = []
    for slot, (name, value) in zip(array, items):
        if not isinstance(value, Value):
            raise TypeError("argument %r must be wrapped with its exact type (u32(...), size_t(...), ptr(...), ...)" % name)
        encoded = name.encode()
        names.append(encoded)
        slot.name = encoded
        slot.type = value.tag
        base = value.tag & 0xff
        if value.tag & T_PTR:
            slot.value.p = value.value
        elif base in _ARRAYS:
            for k, v in enumerate(value.value):
                slot.value.a[k] = v
        elif base in _SIGNED:
            slot.value.i = value.value
        else:
            slot.value.u = value.value
    return array, len(items), names


def _stream_handle(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream  # torch.cuda.Stream


class Kernel:
    """A type-erased kernel wrapper: the analogue of a std::unique_ptr<cuda::registered::kernel_t>."""

    def __init__(self, library, handle):
        self._library, self._handle = library, handle

    def __del__(self):
        if getattr(self, "_handle", None):
            self._library.lib.giddy_kernel_destroy(self._handle)
            self._handle = None

    @property
    def key(self):
        return self._library.lib.giddy_kernel_key(self._handle).decode()

    def get_device_function(self):
        """The address of the wrapped __global__ function (an integer)."""
        return self._library.lib.giddy_kernel_get_device_function(self._handle)

    def attributes(self):
        a = FunctionAttributes()
        self._library.check(self._library.lib.giddy_function_attributes(self.get_device_function(), ctypes.byref(a)))
        return a

    def resolve_launch_configuration(self, extra_arguments, limits=None, device_properties=None, kernel_function_attributes=None):
        """limits: dict with any of block_size, grid_size, dynamic_shared_memory.
        device_properties / kernel_function_attributes: None means "ask the current device / the runtime"."""
        array, count, _names = _marshal(extra_arguments)
        c_limits = None
        if limits:
            c_limits = _Limits()
            for field in ("block_size", "grid_size", "dynamic_shared_memory"):
                if limits.get(field) is not None:
                    setattr(c_limits, "has_" + field, 1)
                    setattr(c_limits, field, int(limits[field]))
        out = LaunchConfig()
        self._library.check(self._library.lib.giddy_kernel_resolve_launch_configuration(
            self._handle,
            ctypes.byref(device_properties) if device_properties is not None else None,
            ctypes.byref(kernel_function_attributes) if kernel_function_attributes is not None else None,
            array, count, ctypes.byref(c_limits) if c_limits is not None else None, ctypes.byref(out)))
        return out

    def enqueue_launch(self, stream, launch_config, arguments):
        """stream: None (torch's current stream), a torch.cuda.Stream, or a raw cudaStream_t value."""
        array, count, _names = _marshal(arguments)
        self._library.check(self._library.lib.giddy_kernel_enqueue_launch(
            self._handle, _stream_handle(stream), ctypes.byref(launch_config), array, count))


class Scan:
    """A host-resident scan (include/giddy_b200.h, group 4): a replayable plan of pinned host -> device chunk copies
    and decode launches on side streams.  Built once, replayed with run().

        scan = library.create_scan(num_streams=4)
        scan.add_copy(chunk, device_tensor[a:b], pinned_host_tensor[a:b])
        scan.add_launch(chunk, kernel, config, {"decompressed": ptr(...), ...})
        scan.add_dependency(chunk, chunk - 1)
        scan.run()            # asynchronous, on torch's current stream
    """

    def __init__(self, library, num_streams):
        self._library = library
        handle = ctypes.c_void_p()
        library.check(library.lib.giddy_scan_create(int(num_streams), ctypes.byref(handle)))
        self._handle = handle
        self._keepalive = []

    def __del__(self):
        if getattr(self, "_handle", None):
            self._library.lib.giddy_scan_destroy(self._handle)
            self._handle = None

    @property
    def num_streams(self):
        return self._library.lib.giddy_scan_num_streams(self._handle)

    @property
    def num_chunks(self):
        return self._library.lib.giddy_scan_num_chunks(self._handle)

    @property
    def num_operations(self):
        return self._library.lib.giddy_scan_num_operations(self._handle)

    def chunk(self, index):
        return ScanChunk(self, index)

    def add_copy(self, chunk, destination, source, nbytes=None):
        """destination: device tensor or address; source: (pinned) host tensor, numpy array or address."""
        if nbytes is None:
            nbytes = source.numel() * source.element_size() if hasattr(source, "numel") else source.nbytes
            if hasattr(destination, "numel") and destination.numel() * destination.element_size() != nbytes:
                raise ValueError("source and destination of a scan copy differ in size")
        self._keepalive.append((destination, source))
        self._library.check(self._library.lib.giddy_scan_add_copy(self._handle, int(chunk), _address_of(destination), _address_of(source), int(nbytes)))

    def add_copy_back(self, chunk, destination, source, nbytes=None):
        """destination: (pinned) host tensor or address; source: device tensor or address."""
        if nbytes is None:
            nbytes = source.numel() * source.element_size()
        self._keepalive.append((destination, source))
        self._library.check(self._library.lib.giddy_scan_add_copy_back(self._handle, int(chunk), _address_of(destination), _address_of(source), int(nbytes)))

    def add_launch(self, chunk, kernel, launch_config, arguments):
        array, count, names = _marshal(arguments)
        self._keepalive.append((kernel, [v.keepalive for v in arguments.values()]))
        self._library.check(self._library.lib.giddy_scan_add_launch(self._handle, int(chunk), kernel._handle, ctypes.byref(launch_config), array, count))

    def add_callback(self, chunk, enqueue):
        """enqueue(stream_handle) is called at every replay and must only enqueue asynchronous work on that stream."""
        def trampoline(stream, _user_data):
            try:
                enqueue(stream)
                return 0
            except Exception:
                return 7
        callback = SCAN_CALLBACK(trampoline)
        self._keepalive.append((callback, enqueue))
        self._library.check(self._library.lib.giddy_scan_add_callback(self._handle, int(chunk), callback, None, None))

    def add_dependency(self, chunk, on_chunk):
        self._library.check(self._library.lib.giddy_scan_add_dependency(self._handle, int(chunk), int(on_chunk)))

    def set_graph_replay(self, enabled=True):
        self._library.check(self._library.lib.giddy_scan_set_graph_replay(self._handle, 1 if enabled else 0))

    def run(self, stream=None):
        self._library.check(self._library.lib.giddy_scan_run(self._handle, _stream_handle(stream)))


class ScanChunk:
    """One chunk of a scan under construction.  Passed as the `stream` of a libgiddy_b200.schemes launch helper, it
    records the launch in the scan instead of enqueueing it."""
    __slots__ = ("scan", "index")

    def __init__(self, scan, index):
        self.scan, self.index = scan, int(index)

    def copy(self, destination, source, nbytes=None):
        self.scan.add_copy(self.index, destination, source, nbytes)

    def copy_back(self, destination, source, nbytes=None):
        self.scan.add_copy_back(self.index, destination, source, nbytes)

    def after(self, on_chunk):
        self.scan.add_dependency(self.index, on_chunk.index if isinstance(on_chunk, ScanChunk) else on_chunk)


class Library:
    """One loaded shared library exporting the registry API (the product, or the rebuilt reference)."""

    def __init__(self, path=_LIB_PATH):
        self.path = path
        self.lib = load_library(path)

    def check(self, status):
        if status != 0:
            message = self.lib.giddy_last_error().decode(errors="replace")
            raise _STATUS_TO_ERROR.get(status, GiddyError)(message)

    def version(self):
        return self.lib.giddy_version().decode()

    def registered_keys(self):
        return [self.lib.giddy_registered_kernel_key(i).decode() for i in range(self.lib.giddy_num_registered_kernels())]

    def can_produce(self, key):
        return bool(self.lib.giddy_can_produce(key.encode()))

    def produce_subclass(self, key):
        handle = ctypes.c_void_p()
        self.check(self.lib.giddy_produce_subclass(key.encode(), ctypes.byref(handle)))
        return Kernel(self, handle)

    def create_scan(self, num_streams=4):
        return Scan(self, num_streams)

    # ---- group 5: shard planning --------------------------------------------------------------------------------

    def split_segments(self, num_segments, parts, part):
        first, past_last = ctypes.c_uint64(), ctypes.c_uint64()
        self.check(self.lib.giddy_split_segments(int(num_segments), int(parts), int(part), ctypes.byref(first), ctypes.byref(past_last)))
        return first.value, past_last.value

    def shard_plan(self, layout, first_segment, past_last_segment):
        shard = Shard()
        self.check(self.lib.giddy_shard_plan(ctypes.byref(layout), int(first_segment), int(past_last_segment), ctypes.byref(shard)))
        return shard

    def slice_bytes(self, piece):
        return self.lib.giddy_slice_bytes(ctypes.byref(piece))

    def slice_copy(self, piece, destination, whole_array):
        """destination, whole_array: host addresses, numpy arrays or CPU torch tensors"""
        self.check(self.lib.giddy_slice_copy(ctypes.byref(piece), _address_of(destination), _address_of(whole_array)))

    def shard_model_coefficients(self, coefficients, coefficient_size, first_element):
        values = (ctypes.c_int64 * len(coefficients))(*[int(c) for c in coefficients])
        out = (ctypes.c_int64 * len(coefficients))()
        self.check(self.lib.giddy_shard_model_coefficients(values, len(coefficients), int(coefficient_size), int(first_element), out))
        return list(out)

    # ---- group 6: host memory ------------------------------------------------------------------------------------

    def host_numa_node_of_device(self, device=-1):
        node = self.lib.giddy_host_numa_node_of_device(int(device))
        return None if node < 0 else node

    def host_bind_thread_to_device(self, device=-1):
        node = self.lib.giddy_host_bind_thread_to_device(int(device))
        return None if node < 0 else node

    def host_alloc(self, nbytes, device=-1):
        """A HostBuffer: page-locked host memory on the device's NUMA node (where the platform has one)."""
        return HostBuffer(self, nbytes, device)

    # ---- group 7: device groups ----------------------------------------------------------------------------------

    def create_device_group(self, devices=None):
        return DeviceGroup(self, devices)

    def device_properties(self, device_id=-1):
        p = DeviceProperties()
        self.check(self.lib.giddy_device_properties(device_id, ctypes.byref(p)))
        return p


class HostBuffer:
    """giddy_host_alloc'd memory, viewable as a numpy array or a (CPU) torch tensor without copying."""

    def __init__(self, library, nbytes, device=-1):
        self._library, self.nbytes = library, int(nbytes)
        pointer, node = ctypes.c_void_p(), ctypes.c_int(-1)
        library.check(library.lib.giddy_host_alloc(self.nbytes, int(device), ctypes.byref(pointer), ctypes.byref(node)))
        self.address = pointer.value or 0
        self.numa_node = None if node.value < 0 else node.value

    def numpy(self, dtype="uint8", offset=0, count=None):
        import numpy as np
        dtype = np.dtype(dtype)
        count = (self.nbytes - offset) // dtype.itemsize if count is None else int(count)
        if count == 0:
            return np.empty(0, dtype)
        buffer = (ctypes.c_ubyte * (count * dtype.itemsize)).from_address(self.address + offset)
        array = np.frombuffer(buffer, dtype=dtype, count=count)
        array.flags.writeable = True
        return array

    def tensor(self, dtype="uint8", offset=0, count=None):
        import torch
        array = self.numpy(dtype, offset, count)
        if array.dtype.kind == "u" and array.dtype.itemsize > 1:
            array = array.view("int%d" % (8 * array.dtype.itemsize))
        return torch.from_numpy(array)

    def free(self):
        if self.address:
            self._library.check(self._library.lib.giddy_host_free(self.address))
            self.address = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceGroup:
    """giddy_device_group_t: one worker thread per GPU; run(work) calls work(rank, size, device) on all of them at once."""

    def __init__(self, library, devices=None):
        self._library = library
        handle = ctypes.c_void_p()
        if devices is None:
            library.check(library.lib.giddy_device_group_create(None, 0, ctypes.byref(handle)))
        else:
            array = (ctypes.c_int * len(devices))(*devices)
            library.check(library.lib.giddy_device_group_create(array, len(devices), ctypes.byref(handle)))
        self._handle = handle

    def __del__(self):
        if getattr(self, "_handle", None):
            self._library.lib.giddy_device_group_destroy(self._handle)
            self._handle = None

    def __len__(self):
        return self._library.lib.giddy_device_group_size(self._handle)

    def device(self, rank):
        return self._library.lib.giddy_device_group_device(self._handle, rank)

    def run(self, work):
        errors = []

        def trampoline(rank, size, device, _user_data):
            try:
                work(rank, size, device)
                return 0
            except Exception as error:  # carried back to the caller's thread
                errors.append(error)
                return 7
        callback = DEVICE_WORK(trampoline)
        status = self._library.lib.giddy_device_group_run(self._handle, callback, None)
        if errors:
            raise errors[0]
        self._library.check(status)

    def broadcast(self, buffers, nbytes, root=0, streams=None):
        """buffers: one device tensor / address per rank; the root's contents reach all the others (NCCL)."""
        addresses = (ctypes.c_void_p * len(buffers))(*[_address_of(b) for b in buffers])
        stream_array = None
        if streams is not None:
            stream_array = (ctypes.c_void_p * len(streams))(*[_stream_handle(s) for s in streams])
        self._library.check(self._library.lib.giddy_device_group_broadcast(self._handle, addresses, int(nbytes), int(root), stream_array))
