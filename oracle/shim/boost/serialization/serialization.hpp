// shim: the reference's DBoW2 headers only name boost::serialization for a member template that is never
// instantiated in the oracle build (no archive type exists here).  TEST INFRASTRUCTURE, NOT PRODUCT.
#ifndef ORBX_SHIM_BOOST_SERIALIZATION_HPP
#define ORBX_SHIM_BOOST_SERIALIZATION_HPP
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> inline Base &base_object(Derived &d) { return static_cast<Base &>(d); }
}}  // namespace boost::serialization
#endif
