// Glue (ours) around the REFERENCE'S OWN decomposition writer, fldecomp/fldgmsh.cpp write_partitions_gmsh (:310-626):
// given a node -> partition map it builds the level-1 / level-2 halos, the trailing-receives node numbering and the
// element order of every partition and writes <name>_<part>.msh + .halo. Compiled together with the unmodified
// reference sources by `make -C oracle ref` into oracle/_ref/libref_fldecomp.so. TEST INFRASTRUCTURE.
#include <deque>
#include <string>
#include <vector>

// prototype as defined in fldecomp/fldgmsh.cpp:310-320 (it has no header declaration)
void write_partitions_gmsh(bool verbose, std::string filename, std::string file_format, const int nparts, const int nnodes,
                           const int dim, const int no_coords, const std::vector<double>& x, const std::vector<int>& decomp,
                           const int nloc, const std::vector<int>& ENList, const std::vector<int>& regionIds, const int snloc,
                           const std::deque<std::vector<int> >& SENList, const std::vector<int>& boundaryIds, int normElemType,
                           int faceType);

// fldgmsh.cpp's decomp_gmsh (not called here) references the METIS front-end of fldecomp/partition.cpp, which is not
// built (METIS is absent). ctypes loads libraries with RTLD_NOW, so the symbols must resolve: they abort if reached.
#include <cstdlib>
#include <set>
namespace Fluidity {
int partition(const std::vector<int>&, const int&, int, int, std::vector<int>&, int, std::vector<int>&) { std::abort(); }
int partition(const std::vector<int>&, int, int, std::vector<int>&, int, std::vector<int>&) { std::abort(); }
int partition(const std::vector<int>&, const std::vector<int>&, const int&, int, int, std::vector<int>&, int, std::vector<int>&) {
  std::abort();
}
int partition(const std::vector<int>&, const std::vector<int>&, int, int, std::vector<int>&, int, std::vector<int>&) { std::abort(); }
}  // namespace Fluidity

extern "C" int ref_write_partitions_gmsh(const char* filename, int nparts, int nnodes, int dim, const double* x,
                                         const int* decomp, int nloc, int nelms, const int* enlist, const int* region_ids,
                                         int snloc, int nfaces, const int* senlist, const int* boundary_ids) {
  std::vector<double> X(x, x + (size_t)nnodes * dim);
  std::vector<int> D(decomp, decomp + nnodes), EN(enlist, enlist + (size_t)nelms * nloc), RI(region_ids, region_ids + nelms);
  std::deque<std::vector<int> > SEN;
  std::vector<int> BI;
  for (int f = 0; f < nfaces; f++) {
    SEN.push_back(std::vector<int>(senlist + (size_t)f * snloc, senlist + (size_t)(f + 1) * snloc));
    BI.push_back(boundary_ids[f]);
  }
  // gmsh element types as decomp_gmsh passes them (fldgmsh.cpp:926-1067): tet 4 / triangle 2 / line 1
  const int normElemType = dim == 3 ? 4 : 2, faceType = dim == 3 ? 2 : 1;
  write_partitions_gmsh(false, filename, "gmsh", nparts, nnodes, dim, dim, X, D, nloc, EN, RI, snloc, SEN, BI, normElemType,
                        faceType);
  return 0;
}
