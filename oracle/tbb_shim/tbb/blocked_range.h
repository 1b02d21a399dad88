// Serial stand-in for <tbb/blocked_range.h> so that the reference's flann/octree.h compiles
// without TBB (absent from this image).  Oracle build only.
#pragma once
namespace tbb {
template <class T>
struct blocked_range {
    T b_, e_;
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
};
}  // namespace tbb
