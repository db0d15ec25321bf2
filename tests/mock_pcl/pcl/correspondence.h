#pragma once
#include <vector>
namespace pcl {
struct Correspondence {
  int index_query, index_match; float distance;
  Correspondence() : index_query(0), index_match(-1), distance(0) {}
  Correspondence(int q, int m, float d) : index_query(q), index_match(m), distance(d) {}
};
typedef std::vector<Correspondence> Correspondences;
}
