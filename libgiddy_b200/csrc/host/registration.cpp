// The factory of kernel wrappers and the registration of the library's 163 wrapper instantiations --
// the same set, under the same keys, as the reference registers in the static_block of each of its
// src/kernel_wrappers/**/*.cu files (cited per group; full key list: SURVEY.md Appendix C).
#include "kernel_wrappers.hpp"
#include "compression_wrappers.hpp"

#include <mutex>

namespace giddy {
namespace host {

// (outside namespace cuda::registered, where the name kernel_t would find the factory's base class)
void register_all_wrappers()
{
	using namespace cuda::kernels;
	namespace models = cuda::functors::unary::parametric_model;
	using ::std::int8_t; using ::std::int16_t; using ::std::int32_t;

	// src/kernel_wrappers/decompression/frame_of_reference.cu:132-142
	{
		using namespace decompression::frame_of_reference;
		kernel_t<4, int16_t, int8_t,  models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int8_t,  models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int16_t, models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<8, int32_t, int16_t, models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<8, int32_t, int16_t, models::linear  <8, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int16_t, int8_t,  models::constant<4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int8_t,  models::constant<4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int16_t, models::constant<4, int32_t>>::registerInSubclassFactory();
	}
	// src/kernel_wrappers/decompression/model.cu:92-96
	{
		using namespace decompression::model;
		kernel_t<4, int16_t, models::linear  <4, int16_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<8, int32_t, models::linear  <8, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int16_t, models::constant<4, int16_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, models::constant<4, int32_t>>::registerInSubclassFactory();
	}
	// src/kernel_wrappers/decompression/run_length_encoding.cu:107-128
	{
		using namespace decompression::run_length_encoding;
		kernel_t<4, 1, 1>::registerInSubclassFactory(); kernel_t<4, 2, 1>::registerInSubclassFactory();
		kernel_t<4, 4, 1>::registerInSubclassFactory(); kernel_t<4, 8, 1>::registerInSubclassFactory();
		kernel_t<4, 1, 2>::registerInSubclassFactory(); kernel_t<4, 2, 2>::registerInSubclassFactory();
		kernel_t<4, 4, 2>::registerInSubclassFactory(); kernel_t<4, 8, 2>::registerInSubclassFactory();
		kernel_t<4, 1, 4>::registerInSubclassFactory(); kernel_t<4, 2, 4>::registerInSubclassFactory();
		kernel_t<4, 4, 4>::registerInSubclassFactory(); kernel_t<4, 8, 4>::registerInSubclassFactory();
		kernel_t<8, 4, 1>::registerInSubclassFactory(); kernel_t<8, 4, 2>::registerInSubclassFactory();
		kernel_t<8, 4, 4>::registerInSubclassFactory();
		kernel_t<8, 1, 8>::registerInSubclassFactory(); kernel_t<8, 4, 8>::registerInSubclassFactory();
		kernel_t<8, 8, 8>::registerInSubclassFactory();
	}
	// src/kernel_wrappers/decompression/run_position_encoding.cu:124-136
	{
		using namespace decompression::run_position_encoding;
		kernel_t<4, 1, 2, true>::registerInSubclassFactory();  kernel_t<4, 2, 2, true>::registerInSubclassFactory();
		kernel_t<4, 4, 2, true>::registerInSubclassFactory();  kernel_t<4, 8, 2, true>::registerInSubclassFactory();
		kernel_t<8, 4, 2, true>::registerInSubclassFactory();
		kernel_t<4, 1, 4, false>::registerInSubclassFactory(); kernel_t<4, 2, 4, false>::registerInSubclassFactory();
		kernel_t<4, 4, 4, false>::registerInSubclassFactory(); kernel_t<4, 8, 4, false>::registerInSubclassFactory();
		kernel_t<8, 4, 8, false>::registerInSubclassFactory();
	}
#define GIDDY_REGISTER_UC_PAIRS(I) \
	kernel_t<I, 2, 1>::registerInSubclassFactory(); kernel_t<I, 4, 1>::registerInSubclassFactory(); \
	kernel_t<I, 4, 2>::registerInSubclassFactory(); kernel_t<I, 8, 1>::registerInSubclassFactory(); \
	kernel_t<I, 8, 2>::registerInSubclassFactory(); kernel_t<I, 8, 4>::registerInSubclassFactory();
	// src/kernel_wrappers/decompression/delta.cu:94-106
	{
		using namespace decompression::delta;
		GIDDY_REGISTER_UC_PAIRS(4) GIDDY_REGISTER_UC_PAIRS(8)
	}
	// src/kernel_wrappers/decompression/discard_zero_bytes_fixed.cu:113-125
	{
		using namespace decompression::discard_zero_bytes::fixed_width;
		GIDDY_REGISTER_UC_PAIRS(4) GIDDY_REGISTER_UC_PAIRS(8)
	}
#undef GIDDY_REGISTER_UC_PAIRS
	// src/kernel_wrappers/decompression/dictionary.cu:113-139
	{
		using namespace decompression::dictionary;
#define GIDDY_REGISTER_DICTIONARY(I) \
		kernel_t<I, 1, 1>::registerInSubclassFactory(); kernel_t<I, 2, 1>::registerInSubclassFactory(); kernel_t<I, 2, 2>::registerInSubclassFactory(); \
		kernel_t<I, 4, 1>::registerInSubclassFactory(); kernel_t<I, 4, 2>::registerInSubclassFactory(); kernel_t<I, 4, 4>::registerInSubclassFactory(); \
		kernel_t<I, 8, 1>::registerInSubclassFactory(); kernel_t<I, 8, 2>::registerInSubclassFactory(); kernel_t<I, 8, 4>::registerInSubclassFactory(); \
		kernel_t<I, 8, 8>::registerInSubclassFactory();
		GIDDY_REGISTER_DICTIONARY(4) GIDDY_REGISTER_DICTIONARY(8)
#undef GIDDY_REGISTER_DICTIONARY
	}
	// src/kernel_wrappers/decompression/discard_zero_bytes_variable.cu:109-119
	{
		using namespace decompression::discard_zero_bytes::variable_width;
		kernel_t<4, 2, 4>::registerInSubclassFactory(); kernel_t<4, 4, 4>::registerInSubclassFactory(); kernel_t<4, 8, 4>::registerInSubclassFactory();
		kernel_t<8, 2, 4>::registerInSubclassFactory(); kernel_t<8, 4, 4>::registerInSubclassFactory(); kernel_t<8, 8, 4>::registerInSubclassFactory();
		kernel_t<4, 2, 8>::registerInSubclassFactory(); kernel_t<4, 4, 8>::registerInSubclassFactory(); kernel_t<4, 8, 8>::registerInSubclassFactory();
	}
	// src/kernel_wrappers/decompression/incidence_bitmaps.cu:96-100
	{
		using namespace decompression::incidence_bitmaps;
		kernel_t<4, 1, 2, true>::registerInSubclassFactory(); kernel_t<4, 2, 2, true>::registerInSubclassFactory();
		kernel_t<4, 4, 2, true>::registerInSubclassFactory(); kernel_t<4, 8, 2, true>::registerInSubclassFactory();
		kernel_t<4, 1, 2, false>::registerInSubclassFactory();
	}
#define GIDDY_REGISTER_SCATTER_E(OI, E) \
	kernel_t<OI, E, 1>::registerInSubclassFactory(); kernel_t<OI, E, 2>::registerInSubclassFactory(); \
	kernel_t<OI, E, 4>::registerInSubclassFactory(); kernel_t<OI, E, 8>::registerInSubclassFactory();
#define GIDDY_REGISTER_SCATTER(OI) \
	GIDDY_REGISTER_SCATTER_E(OI, 1) GIDDY_REGISTER_SCATTER_E(OI, 2) GIDDY_REGISTER_SCATTER_E(OI, 4) GIDDY_REGISTER_SCATTER_E(OI, 8)
	// src/kernel_wrappers/data_layout/scatter.cu:92-124
	{
		using namespace cuda::kernels::scatter;
		GIDDY_REGISTER_SCATTER(4) GIDDY_REGISTER_SCATTER(8)
	}
	// src/kernel_wrappers/data_layout/compressed_indices_scatter.cu:119-151
	{
		using namespace cuda::kernels::scatter::compressed_indices;
		GIDDY_REGISTER_SCATTER(4) GIDDY_REGISTER_SCATTER(8)
	}
#undef GIDDY_REGISTER_SCATTER
#undef GIDDY_REGISTER_SCATTER_E
	// Extensions (no reference counterpart): fused composite decoders
	{
		using namespace decompression::composite::delta_of_discard_zero_bytes_variable_width;
		kernel_t<4, 8, 4, 4>::registerInSubclassFactory(); kernel_t<4, 8, 4, 8>::registerInSubclassFactory();
		kernel_t<4, 8, 2, 4>::registerInSubclassFactory(); kernel_t<4, 4, 2, 4>::registerInSubclassFactory();
		kernel_t<8, 8, 4, 4>::registerInSubclassFactory();
	}
	{
		using namespace decompression::composite::patched_dictionary;
		kernel_t<4, 1, 1>::registerInSubclassFactory(); kernel_t<4, 2, 1>::registerInSubclassFactory();
		kernel_t<4, 4, 1>::registerInSubclassFactory(); kernel_t<4, 8, 1>::registerInSubclassFactory();
		kernel_t<4, 2, 2>::registerInSubclassFactory(); kernel_t<4, 4, 2>::registerInSubclassFactory();
		kernel_t<4, 8, 2>::registerInSubclassFactory(); kernel_t<4, 8, 4>::registerInSubclassFactory();
		kernel_t<8, 8, 1>::registerInSubclassFactory(); kernel_t<8, 4, 2>::registerInSubclassFactory();
	}
	{
		using namespace decompression::composite::patched_frame_of_reference;
		kernel_t<4, int16_t, int8_t,  models::linear<4, int32_t>, 4>::registerInSubclassFactory();
		kernel_t<4, int32_t, int8_t,  models::linear<4, int32_t>, 4>::registerInSubclassFactory();
		kernel_t<4, int32_t, int16_t, models::linear<4, int32_t>, 4>::registerInSubclassFactory();
		kernel_t<8, int32_t, int16_t, models::linear<8, int32_t>, 8>::registerInSubclassFactory();
		kernel_t<4, int16_t, int8_t,  models::constant<4, int32_t>, 4>::registerInSubclassFactory();
		kernel_t<4, int32_t, int8_t,  models::constant<4, int32_t>, 4>::registerInSubclassFactory();
		kernel_t<4, int32_t, int16_t, models::constant<4, int32_t>, 4>::registerInSubclassFactory();
	}
	// Extensions (no reference counterpart: "compression ... is not included", README.md:128-130): the compressors, under the
	// decompressors' type tuples
	{
		using namespace compression::frame_of_reference;
		kernel_t<4, int16_t, int8_t,  models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int8_t,  models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int16_t, models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<8, int32_t, int16_t, models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int16_t, int8_t,  models::constant<4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int8_t,  models::constant<4, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, int16_t, models::constant<4, int32_t>>::registerInSubclassFactory();
	}
#define GIDDY_REGISTER_UC_PAIRS(I) \
	kernel_t<I, 2, 1>::registerInSubclassFactory(); kernel_t<I, 4, 1>::registerInSubclassFactory(); \
	kernel_t<I, 4, 2>::registerInSubclassFactory(); kernel_t<I, 8, 1>::registerInSubclassFactory(); \
	kernel_t<I, 8, 2>::registerInSubclassFactory(); kernel_t<I, 8, 4>::registerInSubclassFactory();
	{ using namespace compression::delta; GIDDY_REGISTER_UC_PAIRS(4) GIDDY_REGISTER_UC_PAIRS(8) }
	{ using namespace compression::discard_zero_bytes::fixed_width; GIDDY_REGISTER_UC_PAIRS(4) GIDDY_REGISTER_UC_PAIRS(8) }
#undef GIDDY_REGISTER_UC_PAIRS
	{
		using namespace compression::dictionary;
#define GIDDY_REGISTER_DICTIONARY(I) \
		kernel_t<I, 1, 1>::registerInSubclassFactory(); kernel_t<I, 2, 1>::registerInSubclassFactory(); kernel_t<I, 2, 2>::registerInSubclassFactory(); \
		kernel_t<I, 4, 1>::registerInSubclassFactory(); kernel_t<I, 4, 2>::registerInSubclassFactory(); kernel_t<I, 4, 4>::registerInSubclassFactory(); \
		kernel_t<I, 8, 1>::registerInSubclassFactory(); kernel_t<I, 8, 2>::registerInSubclassFactory(); kernel_t<I, 8, 4>::registerInSubclassFactory(); \
		kernel_t<I, 8, 8>::registerInSubclassFactory();
		GIDDY_REGISTER_DICTIONARY(4) GIDDY_REGISTER_DICTIONARY(8)
#undef GIDDY_REGISTER_DICTIONARY
	}
	{
		using namespace compression::run_length_encoding;
		kernel_t<4, 1, 1>::registerInSubclassFactory(); kernel_t<4, 2, 1>::registerInSubclassFactory();
		kernel_t<4, 4, 1>::registerInSubclassFactory(); kernel_t<4, 8, 1>::registerInSubclassFactory();
		kernel_t<4, 1, 2>::registerInSubclassFactory(); kernel_t<4, 2, 2>::registerInSubclassFactory();
		kernel_t<4, 4, 2>::registerInSubclassFactory(); kernel_t<4, 8, 2>::registerInSubclassFactory();
		kernel_t<4, 1, 4>::registerInSubclassFactory(); kernel_t<4, 2, 4>::registerInSubclassFactory();
		kernel_t<4, 4, 4>::registerInSubclassFactory(); kernel_t<4, 8, 4>::registerInSubclassFactory();
		kernel_t<8, 4, 1>::registerInSubclassFactory(); kernel_t<8, 4, 2>::registerInSubclassFactory();
		kernel_t<8, 4, 4>::registerInSubclassFactory();
		kernel_t<8, 1, 8>::registerInSubclassFactory(); kernel_t<8, 4, 8>::registerInSubclassFactory();
		kernel_t<8, 8, 8>::registerInSubclassFactory();
	}
	{
		using namespace compression::run_position_encoding;
		kernel_t<4, 1, 2, true>::registerInSubclassFactory();  kernel_t<4, 2, 2, true>::registerInSubclassFactory();
		kernel_t<4, 4, 2, true>::registerInSubclassFactory();  kernel_t<4, 8, 2, true>::registerInSubclassFactory();
		kernel_t<8, 4, 2, true>::registerInSubclassFactory();
		kernel_t<4, 1, 4, false>::registerInSubclassFactory(); kernel_t<4, 2, 4, false>::registerInSubclassFactory();
		kernel_t<4, 4, 4, false>::registerInSubclassFactory(); kernel_t<4, 8, 4, false>::registerInSubclassFactory();
		kernel_t<8, 4, 8, false>::registerInSubclassFactory();
	}
	{
		using namespace compression::discard_zero_bytes::variable_width;
		kernel_t<4, 2, 4>::registerInSubclassFactory(); kernel_t<4, 4, 4>::registerInSubclassFactory(); kernel_t<4, 8, 4>::registerInSubclassFactory();
		kernel_t<8, 2, 4>::registerInSubclassFactory(); kernel_t<8, 4, 4>::registerInSubclassFactory(); kernel_t<8, 8, 4>::registerInSubclassFactory();
		kernel_t<4, 2, 8>::registerInSubclassFactory(); kernel_t<4, 4, 8>::registerInSubclassFactory(); kernel_t<4, 8, 8>::registerInSubclassFactory();
	}
	{
		using namespace compression::incidence_bitmaps;
		kernel_t<4, 1, 2, true>::registerInSubclassFactory(); kernel_t<4, 2, 2, true>::registerInSubclassFactory();
		kernel_t<4, 4, 2, true>::registerInSubclassFactory(); kernel_t<4, 8, 2, true>::registerInSubclassFactory();
		kernel_t<4, 1, 2, false>::registerInSubclassFactory();
	}
	// Utility kernels which the reference has but does not register
	{
		using namespace cuda::kernels::gather;
#define GIDDY_REGISTER_GATHER(I) \
		kernel_t<I, 1, 1>::registerInSubclassFactory(); kernel_t<I, 2, 1>::registerInSubclassFactory(); kernel_t<I, 2, 2>::registerInSubclassFactory(); \
		kernel_t<I, 4, 1>::registerInSubclassFactory(); kernel_t<I, 4, 2>::registerInSubclassFactory(); kernel_t<I, 4, 4>::registerInSubclassFactory(); \
		kernel_t<I, 8, 1>::registerInSubclassFactory(); kernel_t<I, 8, 2>::registerInSubclassFactory(); kernel_t<I, 8, 4>::registerInSubclassFactory(); \
		kernel_t<I, 8, 8>::registerInSubclassFactory();
		GIDDY_REGISTER_GATHER(4) GIDDY_REGISTER_GATHER(8)
#undef GIDDY_REGISTER_GATHER
	}
	{
		using namespace decompression::model::intervals;
		kernel_t<4, int16_t, models::linear  <4, int16_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, models::linear  <4, int32_t>>::registerInSubclassFactory();
		kernel_t<8, int32_t, models::linear  <8, int32_t>>::registerInSubclassFactory();
		kernel_t<4, int16_t, models::constant<4, int16_t>>::registerInSubclassFactory();
		kernel_t<4, int32_t, models::constant<4, int32_t>>::registerInSubclassFactory();
	}
	{
		using namespace cuda::kernels::reduction::scan;
#define GIDDY_REGISTER_SCAN(I) \
		kernel_t<I, 2, 1>::registerInSubclassFactory(); kernel_t<I, 4, 1>::registerInSubclassFactory(); kernel_t<I, 4, 2>::registerInSubclassFactory(); \
		kernel_t<I, 8, 1>::registerInSubclassFactory(); kernel_t<I, 8, 2>::registerInSubclassFactory(); kernel_t<I, 8, 4>::registerInSubclassFactory();
		GIDDY_REGISTER_SCAN(4) GIDDY_REGISTER_SCAN(8)
#undef GIDDY_REGISTER_SCAN
	}
}

} // namespace host
} // namespace giddy

namespace cuda {
namespace registered {

namespace { ::std::once_flag wrappers_registered; }

kernel_t::SubclassFactory& kernel_t::getSubclassFactory_()
{
	static SubclassFactory subclass_factory;
	return subclass_factory;
}

const kernel_t::SubclassFactory& kernel_t::getSubclassFactory()
{
	::std::call_once(wrappers_registered, giddy::host::register_all_wrappers);
	return getSubclassFactory_();
}

::std::unique_ptr<kernel_t> kernel_t::produceSubclass(const factory_key_type& subclass_key)
{
	const auto& factory = getSubclassFactory();
	if (!factory.canProduce(subclass_key)) {
		throw ::std::invalid_argument("No subclass of the base type cuda::registered::kernel_t is registered with key \"" + subclass_key + "\"");
	}
	return factory.produce(subclass_key);
}

bool kernel_t::canProduce(const factory_key_type& subclass_key) { return getSubclassFactory().canProduce(subclass_key); }

} // namespace registered
} // namespace cuda
