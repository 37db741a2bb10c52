// Empty stand-in: C++/include/filter.hpp includes <aruco/aruco.h> but the filter path uses nothing from it.
#ifndef FBUS_REF_STUB_ARUCO
#define FBUS_REF_STUB_ARUCO
namespace aruco {}
#endif
