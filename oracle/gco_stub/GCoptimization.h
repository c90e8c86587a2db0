/* TEST INFRASTRUCTURE ONLY.
 *
 * Recording stand-in for gco-v3.0's GCoptimization.h.  gco is not vendored in the reference tree
 * (reference .gitignore:10, README.md:12-13) and cannot be fetched offline, so the reference's
 * rs_pointcloud_filters.cpp cannot link against the real thing here.  This class has the handful of
 * methods that rspf_smooth_labels calls (rs_pointcloud_filters.cpp:955-971) and simply captures what
 * it is given: that capture is the parity data at the gco *input* boundary (unary data_cost, Potts
 * smooth_cost, initial labels, weighted edges).  swap() is a no-op, whatLabel() echoes the initial
 * label.  The graph-cut result itself is "parity unpinned" (see DESIGN.md).
 */
#ifndef RSGPU_GCO_RECORDING_STUB_H
#define RSGPU_GCO_RECORDING_STUB_H

#include <cassert>
#include <cstring>
#include <cstdint>
#include <vector>

struct gco_capture_t
{
  int n_sites = 0, n_labels = 0;
  std::vector<int> data_cost;    /* n_sites * n_labels */
  std::vector<int> smooth_cost;  /* n_labels * n_labels */
  std::vector<int> init_labels;  /* n_sites */
  std::vector<int> edge_a, edge_b, edge_w;
};

gco_capture_t& gco_last_capture();

class GCoptimizationGeneralGraph
{
public:
  GCoptimizationGeneralGraph( int n_sites, int n_labels )
  {
    gco_capture_t& c = gco_last_capture();
    c = gco_capture_t();
    c.n_sites = n_sites; c.n_labels = n_labels;
    c.init_labels.assign( n_sites, 0 );
  }
  void setDataCost( int* d )
  {
    gco_capture_t& c = gco_last_capture();
    c.data_cost.assign( d, d + (size_t)c.n_sites * c.n_labels );
  }
  void setSmoothCost( int* s )
  {
    gco_capture_t& c = gco_last_capture();
    c.smooth_cost.assign( s, s + (size_t)c.n_labels * c.n_labels );
  }
  void setLabel( int site, int label ) { gco_last_capture().init_labels[site] = label; }
  void setNeighbors( int a, int b, int w )
  {
    gco_capture_t& c = gco_last_capture();
    c.edge_a.push_back( a ); c.edge_b.push_back( b ); c.edge_w.push_back( w );
  }
  void swap( int ) {}
  int whatLabel( int site ) { return gco_last_capture().init_labels[site]; }
};

#endif
