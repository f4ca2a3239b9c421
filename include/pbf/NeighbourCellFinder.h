/* NeighbourCellFinder.h -- shim with the public interface of the reference's NeighbourCellFinder
 * (src/NeighbourCellFinder.h:35-115):
 *   NeighbourCellFinder(const GLuint &numparticles, const glm::ivec3 &gridsize);
 *   void FindNeighbourCells(const GLuint &particlebuffer);   const Texture &GetResult() const;
 * FindNeighbourCells is the reorder + cell start/end + merged-run stage of the owning SPH handle (pbf_build_cells).
 * The reference returns the 48 B/particle run list as a GL buffer texture; the CUDA kernels rebuild the runs from the
 * cell tables on the fly, so GetResult() is replaced by GetRuns() (host read-back of the same 9 runs per particle). */
#ifndef PBF_SHIM_NEIGHBOURCELLFINDER_H
#define PBF_SHIM_NEIGHBOURCELLFINDER_H

#include <vector>

#include "shim_common.h"

class NeighbourCellFinder {
public:
    NeighbourCellFinder(const GLuint &_numparticles, const glm::ivec3 &_gridsize)
        : numparticles(_numparticles), gridsize(_gridsize), handle(nullptr) {}
    void Attach(pbf_handle h) { handle = h; }
    void FindNeighbourCells(const GLuint & /*particlebuffer*/) {
        pbf_shim::check(pbf_build_cells(handle), "NeighbourCellFinder::FindNeighbourCells");
    }
    /* 9 x (start, count) per sorted particle, in the order of neighbourcells.glsl:37-47 */
    void GetRuns(std::vector<int32_t> &run_start, std::vector<int32_t> &run_count) const {
        run_start.resize(9 * (size_t)numparticles);
        run_count.resize(9 * (size_t)numparticles);
        pbf_shim::check(pbf_get_neighbour_runs(handle, run_start.data(), run_count.data()), "NeighbourCellFinder::GetRuns");
    }
    /* gridtexture / gridendtexture (src/NeighbourCellFinder.cpp:59-90) */
    void GetCellRanges(std::vector<int32_t> &start, std::vector<int32_t> &end) const {
        size_t nc = (size_t)gridsize.x * gridsize.y * gridsize.z;
        start.resize(nc);
        end.resize(nc);
        pbf_shim::check(pbf_get_cell_ranges(handle, start.data(), end.data()), "NeighbourCellFinder::GetCellRanges");
    }

private:
    const GLuint numparticles;
    const glm::ivec3 gridsize;
    pbf_handle handle;
};

#endif
