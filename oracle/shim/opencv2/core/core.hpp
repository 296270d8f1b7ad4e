// shim: see ../opencv.hpp
#include "../opencv.hpp"
