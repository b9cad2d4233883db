// Thin C wrapper around the REFERENCE's own Bresenham iterator, compiled from where it lies
// (-I/root/reference/nav2_util/include; see Makefile target `ref`).  No reference source is
// copied into this repository: this file only includes the header and walks the iterator.
// Test infrastructure: used by tests/test_oracle_costmap.py to pin oracle::LineIterator and
// the CUDA Bresenham against the real nav2_util::LineIterator.
#include "nav2_util/line_iterator.hpp"

extern "C" int ref_line_cells(int x0, int y0, int x1, int y1, int * xs, int * ys, int cap)
{
  int n = 0;
  for (nav2_util::LineIterator line(x0, y0, x1, y1); line.isValid(); line.advance()) {
    if (n < cap) {xs[n] = line.getX(); ys[n] = line.getY();}
    ++n;
  }
  return n;
}
