// giddy/registered_wrapper.hpp -- the string-keyed factory of kernel wrappers.
//
// Mirrors cuda::registered::kernel_t (reference src/kernel_wrappers/registered_wrapper.h:28-39) and
// the util::Factory / util::mixins::FactoryProducible machinery behind it (src/util/Factory.h:36-102,
// src/util/FactoryProducible.h:26-95).  The keys are the reference's: the
// demangled wrapper class name without its leading "cuda::kernels::".
//
// Unlike the reference, registration does not depend on static-initialisation order or on
// --whole-archive linking: the factory is filled on first use by the library
// (libgiddy_b200/csrc/host/registration.cpp).
#pragma once
#ifndef GIDDY_REGISTERED_WRAPPER_HPP_
#define GIDDY_REGISTERED_WRAPPER_HPP_

#include "kernel_wrapper.hpp"

#include <memory>
#include <string>
#include <unordered_map>

namespace util {

template <typename Key, typename T>
class Factory {
public:
	using Instantiator  = T* (*)();
	using Instantiators = ::std::unordered_map<Key, Instantiator>;

	template <typename U>
	void registerClass(const Key& key, bool ignore_repeat_registration = false)
	{
		static_assert(::std::is_base_of<T, U>::value, "can only register classes derived from the factory's base class");
		if (instantiators_.find(key) != instantiators_.end()) {
			if (ignore_repeat_registration) { return; }
			throw ::std::logic_error("Repeat registration of the same subclass in this factory.");
		}
		instantiators_.emplace(key, [] () -> T* { return new U(); });
	}

	::std::unique_ptr<T> produce(const Key& key) const
	{
		auto it = instantiators_.find(key);
		if (it == instantiators_.end()) {
			throw ::std::invalid_argument("Attempt to produce an instance of a class by a key not registered in the factory");
		}
		return ::std::unique_ptr<T>(it->second());
	}

	bool canProduce(const Key& key) const { return instantiators_.find(key) != instantiators_.end(); }

	const Instantiators& getInstantiators() const { return instantiators_; }
	const Instantiators& getInsantiators() const { return instantiators_; } // the reference's spelling (Factory.h:98)

protected:
	Instantiators instantiators_;
};

template <typename Key, typename T> using ExposedFactory = Factory<Key, T>;

} // namespace util

namespace cuda {
namespace registered {

class GIDDY_API kernel_t : public cuda::kernel_t {
public:
	using factory_key_type = ::std::string;
	using parent           = cuda::kernel_t;
	using SubclassFactory  = util::ExposedFactory<factory_key_type, kernel_t>;

	/// The factory, with all of the library's wrappers registered.
	static const SubclassFactory& getSubclassFactory();

	/// @throws std::invalid_argument when no wrapper is registered under the key
	static ::std::unique_ptr<kernel_t> produceSubclass(const factory_key_type& subclass_key);

	static bool canProduce(const factory_key_type& subclass_key);

	template <typename U>
	static void registerInSubclassFactory(const factory_key_type& key, bool ignore_repeat_registration = true)
	{
		getSubclassFactory_().template registerClass<U>(key, ignore_repeat_registration);
	}

protected:
	static SubclassFactory& getSubclassFactory_();
};

} // namespace registered
} // namespace cuda

#endif // GIDDY_REGISTERED_WRAPPER_HPP_
