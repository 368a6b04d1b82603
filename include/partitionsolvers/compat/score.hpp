// Forwarding header (see DP.hpp in this directory).
#include <partitionsolvers/score.hpp>
