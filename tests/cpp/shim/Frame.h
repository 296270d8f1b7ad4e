// tests/cpp/shim/Frame.h — lets include/ORBmatcher.h (which includes the reference's "Frame.h") compile in a container without
// the reference's dependency tree: forwards to the Frame / MapPoint stand-in of oracle/shim.  TEST INFRASTRUCTURE.
#include "frame_shim.h"
