// Drop-in for reference image_compression/include/QR.hpp: the QR class hierarchy lives in ../QR.hpp.
#include "../QR.hpp"
