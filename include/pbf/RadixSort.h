/* RadixSort.h -- shim with the public interface of the reference's RadixSort (src/RadixSort.h:33-131):
 *   RadixSort(GLuint blocksize, GLuint numblocks, const glm::ivec3 &gridsize); GLuint GetBuffer() const; void Run();
 * The reference sorts the 16-byte particle records in its own GL buffer, 2 key bits per pass.  Here Run() is the
 * onesweep stage of the owning SPH handle (pbf_sort); the sorted records are read back with GetSortedRecords().
 * The SPH shim constructs it with the reference's arguments and then attaches its handle. */
#ifndef PBF_SHIM_RADIXSORT_H
#define PBF_SHIM_RADIXSORT_H

#include <vector>

#include "shim_common.h"

class RadixSort {
public:
    RadixSort(GLuint _blocksize, GLuint _numblocks, const glm::ivec3 &gridsize)
        : blocksize(_blocksize), numblocks(_numblocks), handle(nullptr) {
        if (blocksize & 1) throw std::logic_error("The block size for sorting has to be even."); /* src/RadixSort.cpp:41-42 */
        int32_t g[3] = {gridsize.x, gridsize.y, gridsize.z};
        numbits = pbf_sort_bits(g); /* = 2 * reference pass count (src/RadixSort.cpp:44, :127) */
    }
    void Attach(pbf_handle h) { handle = h; }
    /* reference: GL name of the current record buffer, valid until the next Run (src/RadixSort.h:45-53).  The records
     * live in CUDA memory here; 0 is returned and NeighbourCellFinder::FindNeighbourCells ignores its argument. */
    GLuint GetBuffer(void) const { return 0; }
    void Run(void) { pbf_shim::check(pbf_sort(handle), "RadixSort::Run"); }
    /* sorted {pos, id} records + keys, host side (debug) */
    void GetSortedRecords(std::vector<float> &records, std::vector<uint32_t> &keys) const {
        uint32_t n = pbf_num_particles(handle);
        records.resize(4 * (size_t)n);
        keys.resize(n);
        pbf_shim::check(pbf_get_sorted(handle, keys.data(), nullptr, records.data()), "RadixSort::GetSortedRecords");
    }
    unsigned int GetNumBits(void) const { return numbits; }

private:
    GLuint blocksize, numblocks;
    unsigned int numbits;
    pbf_handle handle;
};

#endif
