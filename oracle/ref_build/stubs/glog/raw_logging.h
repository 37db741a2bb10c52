// stand-in for <glog/raw_logging.h> (nothing of it is used by the filter path)
#include "logging.h"
