// Serial stand-in for <tbb/parallel_for.h> (see blocked_range.h).  Oracle build only.
#pragma once
#include "blocked_range.h"
namespace tbb {
template <class R, class F>
void parallel_for(const R& r, const F& f) { f(r); }
}  // namespace tbb
