#include "pgo_opencv_shim.h"
