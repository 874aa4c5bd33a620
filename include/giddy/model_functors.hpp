// giddy/model_functors.hpp -- the parametric model types which appear as template arguments of the
// Frame-of-Reference and Model wrappers (and therefore in their factory keys).
//
// Mirrors cuda::functors::unary::parametric_model::{constant, linear} of the reference,
// src/cuda/model_functors.hpp:26-111.  Host-side evaluation is provided for compressors and tests;
// the device side works on the raw coefficient arrays.
#pragma once
#ifndef GIDDY_MODEL_FUNCTORS_HPP_
#define GIDDY_MODEL_FUNCTORS_HPP_

#include "types.hpp"

namespace cuda {
namespace functors {
namespace unary {
namespace parametric_model {

template <unsigned IndexSize, typename T, unsigned ModelDimension>
struct function_t {
	enum { index_size = IndexSize };
	enum : unsigned { model_dimension = ModelDimension };
	using index_type        = util::uint_t<IndexSize>;
	using size_type         = index_type;
	using coefficient_type  = T;
	using argument_type     = index_type;
	using result_type       = T;
	using coefficients_type = cuda::array<coefficient_type, ModelDimension>;

	function_t(const coefficients_type& model_coefficients_) : model_coefficients(model_coefficients_) { }

	const coefficients_type model_coefficients;
};

/// x -> c[0] + c[1] * x, in the usual arithmetic conversions of T and the index type, converted to T
template <unsigned IndexSize, typename T>
struct linear : public function_t<IndexSize, T, 2> {
	using parent = function_t<IndexSize, T, 2>;
	using parent::parent;
	using index_type = typename parent::index_type;
	enum : int { kind = GIDDY_MODEL_LINEAR };
	T operator()(const index_type& x) const { return (T)(this->model_coefficients[0] + this->model_coefficients[1] * x); }
};

/// x -> c[0]
template <unsigned IndexSize, typename T>
struct constant : public function_t<IndexSize, T, 1> {
	using parent = function_t<IndexSize, T, 1>;
	using parent::parent;
	using index_type = typename parent::index_type;
	enum : int { kind = GIDDY_MODEL_CONSTANT };
	T operator()(const index_type&) const { return this->model_coefficients[0]; }
};

} // namespace parametric_model
} // namespace unary
} // namespace functors
} // namespace cuda

#endif // GIDDY_MODEL_FUNCTORS_HPP_
