// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in. The reference includes
// <boost/container/set.hpp> (include/index_bipartite.h:1) but never uses anything from it.
#pragma once
