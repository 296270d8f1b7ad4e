// shim: see serialization.hpp
#include "serialization.hpp"
