/* Stand-in for gco-v3.0's GCoptimization.h, used ONLY to link the segment_transfer drop-in test executables in this
 * repository.  gco is not vendored in the reference tree (its licence forbids redistribution, reference README.md:12-13)
 * and cannot be fetched offline.  The class has the methods rspf_smooth_labels calls (lib/rs/rs_pointcloud_filters.cpp:
 * 955-971); swap() does nothing and whatLabel() returns the initial label, so the executables built with it skip the
 * graph cut: both the CPU reference build and the rsgpu build, which is what the drop-in test compares.  A Rescan
 * maintainer builds against the real lib/gco instead. */
#ifndef RSGPU_GCO_PASSTHROUGH_H
#define RSGPU_GCO_PASSTHROUGH_H

#include <cassert>
#include <cstring>
#include <cstdint>
#include <vector>

class GCoptimizationGeneralGraph
{
public:
  GCoptimizationGeneralGraph( int n_sites, int n_labels ) : n_sites_( n_sites ), n_labels_( n_labels ), labels_( n_sites > 0 ? n_sites : 0, 0 ) {}
  void setDataCost( int* ) {}
  void setSmoothCost( int* ) {}
  void setLabel( int site, int label ) { labels_[site] = label; }
  void setNeighbors( int, int, int ) {}
  void swap( int ) {}
  int whatLabel( int site ) { return labels_[site]; }

private:
  int n_sites_, n_labels_;
  std::vector<int> labels_;
};

#endif
