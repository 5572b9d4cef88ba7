/* TEST INFRASTRUCTURE: just enough of boost::serialization's names for DBoW2's FeatureVector.h / BowVector.h to parse. */
#ifndef DVM_SLAMSHIM_BOOST_SERIALIZATION
#define DVM_SLAMSHIM_BOOST_SERIALIZATION
namespace boost { namespace serialization {
class access;
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
} }
#endif
