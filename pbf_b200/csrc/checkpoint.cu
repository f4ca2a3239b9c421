// checkpoint.cu -- state files: dump / resume of a simulation.
//
// The reference keeps its whole state in three GL buffers (position, velocity, highlight by particle id,
// src/SPH.cpp:106-133) plus the parameter block (src/SPH.h:252-285) and has no way to save it; everything else the step
// touches (sorted records, cell tables, lambda, vorticity) is rebuilt from those buffers by every SPH::Run.  A state
// file therefore holds exactly that, and a run resumed from one continues bit-for-bit like the run that wrote it (the
// step is a pure function of this state: stable sort, no floating-point atomics).
//
// Layout (little endian, no padding between sections):
//   header   128 bytes  magic "PBFB200S", version, header size, N, grid, wall, ref_quirks, pbf_params, step counter,
//                       FNV-1a-64 of the payload, pbf_options
//   payload  N x float4 positions, N x float4 velocities, N x uint32 highlight flags
// The *_file functions work on HOST arrays and need no device; pbf_save_state / pbf_load_state wrap them with the
// handle's download / upload.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "pbf_internal.cuh"

namespace {

const char MAGIC[8] = {'P', 'B', 'F', 'B', '2', '0', '0', 'S'};
constexpr uint32_t VERSION = 1;

struct Header {                 // 128 bytes
    char magic[8];
    uint32_t version, header_bytes;
    uint64_t num_particles;
    int32_t grid[3];
    float wall[3];
    int32_t ref_quirks;
    pbf_params params;          // 44 bytes, ends at offset 96
    uint64_t steps;
    uint64_t checksum;
    int32_t self_term;          // pbf_options; all-zero (files written before the options existed) = everything off
    int32_t has_restitution;
    float restitution;
    int32_t full_support;       // was reserved (zero) in files written before the option existed
};
static_assert(sizeof(Header) == 128, "state file header is 128 bytes");

int fail(int code, const std::string &msg) {
    pbf_set_error(msg);
    return code;
}

uint64_t fnv1a(uint64_t h, const void *data, size_t n) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    // eight bytes per round keeps this far below disk speed; the order of bytes inside the word is the file's
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 1099511628211ull;
    }
    for (; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

uint64_t payload_sum(const float *pos4, const float *vel4, const uint32_t *hl, size_t n) {
    uint64_t h = 14695981039346656037ull;
    h = fnv1a(h, pos4, n * 16);
    h = fnv1a(h, vel4, n * 16);
    h = fnv1a(h, hl, n * 4);
    return h;
}

int read_header(FILE *f, const char *path, Header *hd) {
    if (fread(hd, 1, sizeof(*hd), f) != sizeof(*hd)) return fail(PBF_ERR_INVALID, std::string(path) + ": truncated header");
    if (memcmp(hd->magic, MAGIC, 8) != 0) return fail(PBF_ERR_INVALID, std::string(path) + ": not a pbf_b200 state file");
    if (hd->version != VERSION || hd->header_bytes != sizeof(Header))
        return fail(PBF_ERR_INVALID, std::string(path) + ": unsupported state file version");
    if (hd->num_particles == 0 || hd->num_particles >= (1ull << 30))
        return fail(PBF_ERR_INVALID, std::string(path) + ": implausible particle count");
    return PBF_OK;
}

void to_info(const Header &hd, pbf_state_info *info) {
    info->num_particles = (uint32_t)hd.num_particles;
    for (int a = 0; a < 3; a++) { info->grid[a] = hd.grid[a]; info->wall[a] = hd.wall[a]; }
    info->ref_quirks = hd.ref_quirks;
    info->params = hd.params;
    info->steps = hd.steps;
    info->options.density_self_term = hd.self_term;
    info->options.wall_restitution = hd.has_restitution ? hd.restitution : -1.0f;
    info->options.full_support = hd.full_support;
}

}  // namespace

extern "C" {

int pbf_state_file_write(const char *path, const pbf_state_info *info, const float *pos4, const float *vel4,
                         const uint32_t *highlight) {
    if (!path || !info || !pos4) return fail(PBF_ERR_INVALID, "pbf_state_file_write: null argument");
    const size_t n = info->num_particles;
    if (n == 0) return fail(PBF_ERR_INVALID, "pbf_state_file_write: no particles");
    std::vector<float> zv;
    std::vector<uint32_t> zh;
    if (!vel4) { zv.assign(n * 4, 0.0f); vel4 = zv.data(); }
    if (!highlight) { zh.assign(n, 0u); highlight = zh.data(); }
    Header hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, MAGIC, 8);
    hd.version = VERSION;
    hd.header_bytes = sizeof(Header);
    hd.num_particles = n;
    for (int a = 0; a < 3; a++) { hd.grid[a] = info->grid[a]; hd.wall[a] = info->wall[a]; }
    hd.ref_quirks = info->ref_quirks;
    hd.params = info->params;
    hd.steps = info->steps;
    hd.self_term = info->options.density_self_term;
    hd.has_restitution = info->options.wall_restitution >= 0.0f ? 1 : 0;
    hd.restitution = hd.has_restitution ? info->options.wall_restitution : 0.0f;
    hd.full_support = info->options.full_support;
    hd.checksum = payload_sum(pos4, vel4, highlight, n);
    // write next to the target and rename: a crash never leaves a half-written file under the final name
    const std::string tmp = std::string(path) + ".part";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(PBF_ERR_INVALID, "pbf_state_file_write: cannot open " + tmp);
    bool ok = fwrite(&hd, 1, sizeof(hd), f) == sizeof(hd) && fwrite(pos4, 16, n, f) == n && fwrite(vel4, 16, n, f) == n &&
              fwrite(highlight, 4, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return fail(PBF_ERR_INVALID, std::string("pbf_state_file_write: write to ") + path + " failed");
    }
    return PBF_OK;
}

int pbf_state_file_info(const char *path, pbf_state_info *info) {
    if (!path || !info) return fail(PBF_ERR_INVALID, "pbf_state_file_info: null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(PBF_ERR_INVALID, std::string("pbf_state_file_info: cannot open ") + path);
    Header hd;
    const int r = read_header(f, path, &hd);
    fclose(f);
    if (r) return r;
    to_info(hd, info);
    return PBF_OK;
}

int pbf_state_file_read(const char *path, pbf_state_info *info, float *pos4, float *vel4, uint32_t *highlight,
                        uint32_t capacity) {
    if (!path || !pos4 || !vel4 || !highlight) return fail(PBF_ERR_INVALID, "pbf_state_file_read: null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(PBF_ERR_INVALID, std::string("pbf_state_file_read: cannot open ") + path);
    Header hd;
    int r = read_header(f, path, &hd);
    if (r) { fclose(f); return r; }
    const size_t n = (size_t)hd.num_particles;
    if (n > capacity) { fclose(f); return fail(PBF_ERR_CAPACITY, std::string(path) + ": more particles than the caller's arrays hold"); }
    const bool ok = fread(pos4, 16, n, f) == n && fread(vel4, 16, n, f) == n && fread(highlight, 4, n, f) == n;
    fclose(f);
    if (!ok) return fail(PBF_ERR_INVALID, std::string(path) + ": truncated payload");
    if (payload_sum(pos4, vel4, highlight, n) != hd.checksum) return fail(PBF_ERR_INVALID, std::string(path) + ": checksum mismatch");
    if (info) to_info(hd, info);
    return PBF_OK;
}

int pbf_save_state(pbf_handle s, const char *path) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    if (!path) return fail(PBF_ERR_INVALID, "pbf_save_state: null path");
    if (s->slab) return fail(PBF_ERR_STATE, "pbf_save_state: slab handles are saved through pbf_slab_download by the host runtime");
    std::vector<float> pos((size_t)s->n * 4), vel((size_t)s->n * 4);
    std::vector<uint32_t> hl(s->n);
    const int r = pbf_download_state(s, pos.data(), vel.data(), hl.data());
    if (r) return r;
    pbf_state_info info;
    info.num_particles = s->n;
    for (int a = 0; a < 3; a++) { info.grid[a] = s->cfg.grid[a]; info.wall[a] = s->cfg.wall[a]; }
    info.ref_quirks = s->cfg.ref_quirks;
    info.params = s->params;
    info.options = s->options;
    info.steps = s->steps;
    return pbf_state_file_write(path, &info, pos.data(), vel.data(), hl.data());
}

int pbf_load_state(pbf_handle s, const char *path) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    if (!path) return fail(PBF_ERR_INVALID, "pbf_load_state: null path");
    if (s->slab) return fail(PBF_ERR_STATE, "pbf_load_state: slab handles are loaded through pbf_slab_upload by the host runtime");
    pbf_state_info info;
    int r = pbf_state_file_info(path, &info);
    if (r) return r;
    if (info.num_particles != s->n || info.grid[0] != s->cfg.grid[0] || info.grid[1] != s->cfg.grid[1] ||
        info.grid[2] != s->cfg.grid[2])
        return fail(PBF_ERR_INVALID, std::string(path) + ": particle count or grid differs from the handle's (create the handle "
                                                         "from pbf_state_file_info)");
    // walls and the findcells quirk mode are constructor arguments like the grid: a run resumed under different ones would
    // not continue bit for bit
    if (info.wall[0] != s->cfg.wall[0] || info.wall[1] != s->cfg.wall[1] || info.wall[2] != s->cfg.wall[2] ||
        (info.ref_quirks != 0) != (s->cfg.ref_quirks != 0))
        return fail(PBF_ERR_INVALID, std::string(path) + ": wall offsets or ref_quirks differ from the handle's (create the "
                                                         "handle from pbf_state_file_info)");
    std::vector<float> pos((size_t)s->n * 4), vel((size_t)s->n * 4);
    std::vector<uint32_t> hl(s->n);
    r = pbf_state_file_read(path, &info, pos.data(), vel.data(), hl.data(), s->n);
    if (r) return r;
    r = pbf_upload_state(s, pos.data(), vel.data(), s->n);      // clears the highlight buffer (src/Simulation.cpp:271-272)
    if (r) return r;
    {
        DeviceGuard guard(s->device);
        GlScope gl(s);
        if (gl.rc) return gl.rc;
        PBF_CUDA(cudaMemcpyAsync(s->hl, hl.data(), (size_t)s->n * 4, cudaMemcpyHostToDevice, s->stream));
        PBF_CUDA(cudaStreamSynchronize(s->stream));
    }
    r = pbf_set_params(s, &info.params);
    if (r) return r;
    r = pbf_set_options(s, &info.options);
    if (r) return r;
    s->steps = info.steps;
    return PBF_OK;
}

uint64_t pbf_step_count(pbf_handle s) { return s ? s->steps : 0; }

}  // extern "C"
