// Forwarding header: lets sources written against the reference (`#include "DP.hpp"`) pick up the B200-backed
// facade.  A maintainer replaces src/cpp/include/{DP,DP_impl,score,score_impl,LTSS,LTSS_impl}.hpp with the three
// files of this directory (quote-includes resolve next to the including file first, so shadowing through -I alone
// is not enough).  The standard headers the reference's DP.hpp pulls in are kept, because its callers rely on them.
#include <algorithm>
#include <cmath>
#include <exception>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <utility>
#include <vector>

#include <partitionsolvers/DP.hpp>
#include <partitionsolvers/LTSS.hpp>
#include <partitionsolvers/score.hpp>
using namespace Objectives;
