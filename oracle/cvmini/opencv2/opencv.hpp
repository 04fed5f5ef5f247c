// Forwarding header of the OpenCV stand-in (see oracle/cvmini/cvmini.hpp).  Test infrastructure only.
#include "cvmini.hpp"
