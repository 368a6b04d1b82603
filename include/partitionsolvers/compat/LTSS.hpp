// Forwarding header (see DP.hpp in this directory).
#include <algorithm>
#include <cmath>
#include <exception>
#include <iostream>
#include <iterator>
#include <limits>
#include <list>
#include <memory>
#include <utility>
#include <vector>

#include <partitionsolvers/LTSS.hpp>
#include <partitionsolvers/score.hpp>
using namespace Objectives;
