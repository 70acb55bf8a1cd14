// oracle/shim: see ros/ros.h (TEST INFRASTRUCTURE ONLY).
#include "ros.h"
