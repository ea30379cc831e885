// multi.cu — one call, all GPUs of the box: contiguous read ranges per device, NUMA-local host memory.
//
// The reference fans a batch out over all its workers in one call and keeps the order of the rows
// (composition/src/oligo.rs:126-143: buffer.par_iter().map(vectorise_one).collect(); pybindings/src/oligo.rs:77-81:
// seqs.into_par_iter()).  The analogue here: ktb_multi_vectorise cuts the batch into one contiguous range of
// sequences per device, balanced by bases (the cut points are where the prefix sum of lengths crosses r/G of the
// total), and runs ktb_oligo_vectorise on every range from its own host thread — one handle, three stream / buffer
// sets per device — each device writing its own slab of rows.  Rows are independent, so there is no collective.
//
// Host memory: the D2H traffic of 8 devices (8 x 20 GB per step in the headline config) must not land on one NUMA
// node.  ktb_host_alloc_near() page-locks memory that is bound to the NUMA node of a GPU (mmap + mbind +
// cudaHostRegister; the node comes from /sys/bus/pci/devices/<bdf>/numa_node, not from the — possibly cgroup-clipped —
// CPU affinity mask), ktb_multi_alloc_rows() does the same slab by slab for the output of a multi-device call.
#include "../../include/kmertools_b200.h"
#include "device_guard.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

int ktb_internal_fail(int code, const char *msg);

namespace {

// ---- NUMA-aware pinned allocations ---------------------------------------------------------------------------------
struct HostBlock {
    size_t bytes = 0;
    bool mapped = false;   // mmap + cudaHostRegister (else cudaHostAlloc)
};
std::mutex g_host_mutex;
std::map<void *, HostBlock> g_host_blocks;

int numa_node_of_device(int device) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/numa_node";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// mbind(2) without libnuma: MPOL_BIND = 2 would fail the allocation when the node is not allowed (cgroup cpuset.mems);
// MPOL_PREFERRED = 1 falls back to another node instead, which is what a library should do.
bool bind_range_to_node(void *p, size_t bytes, int node) {
    if (node < 0 || node >= 1024) return false;
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    return syscall(SYS_mbind, p, bytes, 1 /* MPOL_PREFERRED */, mask, sizeof(mask) * 8, 0) == 0;
}

// One anonymous mapping, part i ([cuts[i], cuts[i+1]) rounded to pages) preferred on node[i], touched, page-locked.
void *alloc_mapped(size_t bytes, const std::vector<size_t> &cuts, const std::vector<int> &nodes) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = ((bytes ? bytes : 1) + page - 1) & ~(page - 1);
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    for (size_t i = 0; i + 1 < cuts.size(); ++i) {
        const size_t a = std::min(len, (cuts[i] + page - 1) & ~(page - 1));   // a page belongs to the part that starts in it
        const size_t b = std::min(len, (cuts[i + 1] + page - 1) & ~(page - 1));
        if (b > a) bind_range_to_node((char *)p + a, b - a, nodes[i]);
    }
    // first touch under the policy (in parallel: page faults of tens of GB are slow from one thread)
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=] {
            const size_t a = (len / page) * t / nt * page, b = (len / page) * (t + 1) / nt * page;
            for (size_t o = a; o < b; o += page) ((volatile char *)p)[o] = 0;
        });
    for (auto &x : th) x.join();
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, len);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_host_mutex);
    g_host_blocks[p] = HostBlock{len, true};
    return p;
}

}  // namespace

// the registry also knows the plain cudaHostAlloc blocks of ktb_host_alloc (api.cu), so ktb_host_free frees both kinds
__attribute__((visibility("hidden"))) void ktb_internal_register_host(void *p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_host_mutex);
    g_host_blocks[p] = HostBlock{bytes, false};
}
__attribute__((visibility("hidden"))) void ktb_internal_free_host(void *p) {
    HostBlock b;
    {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        auto it = g_host_blocks.find(p);
        if (it == g_host_blocks.end()) {   // not ours (or allocated before the registry existed): assume cudaHostAlloc
            cudaFreeHost(p);
            return;
        }
        b = it->second;
        g_host_blocks.erase(it);
    }
    if (b.mapped) {
        cudaHostUnregister(p);
        munmap(p, b.bytes);
    } else {
        cudaFreeHost(p);
    }
}

struct ktb_multi {
    int k = 0;
    std::vector<int> devices;
    std::vector<ktb_oligo *> handles;
    std::vector<ktb_stats> stats;
    std::vector<uint64_t> bounds;   // of the last call
};

extern "C" {

int ktb_device_numa_node(int device) {
    if (device < 0 || device >= ktb_device_count()) return -1;
    return numa_node_of_device(device);
}

void *ktb_host_alloc_near(size_t bytes, int device) {
    const int node = ktb_device_numa_node(device);
    if (node >= 0) {
        void *p = alloc_mapped(bytes, {0, bytes}, {node});
        if (p) return p;
    }
    return ktb_host_alloc(bytes);   // unknown topology or mapping refused: plain page-locked memory
}

int ktb_shard_bounds(const uint64_t *offsets, uint64_t n, int parts, uint64_t *bounds) {
    if (!offsets || !bounds || parts < 1) return ktb_internal_fail(KTB_ERR_ARG, "bad argument");
    const uint64_t base0 = offsets[0], total = offsets[n] - offsets[0];
    bounds[0] = 0;
    for (int r = 1; r < parts; ++r) {
        uint64_t cut;
        if (total == 0) {   // nothing but empty sequences: split by count
            cut = n * (uint64_t)r / (uint64_t)parts;
        } else {            // first sequence that starts at or after r/parts of the bases
            const uint64_t target = base0 + (uint64_t)(((unsigned __int128)total * (unsigned)r) / (unsigned)parts);
            cut = (uint64_t)(std::lower_bound(offsets, offsets + n + 1, target) - offsets);
        }
        cut = std::min(cut, n);
        bounds[r] = std::max(cut, bounds[r - 1]);
    }
    bounds[parts] = n;
    return KTB_OK;
}

int ktb_multi_create(int k, const int *devices, int ndev, ktb_multi **out) {
    if (!out) return ktb_internal_fail(KTB_ERR_ARG, "out is NULL");
    *out = nullptr;
    const int visible = ktb_device_count();
    if (visible <= 0) return ktb_internal_fail(KTB_ERR_NODEVICE, "no CUDA device available (this library has no CPU path)");
    if (ndev < 0 || (ndev > 0 && !devices)) return ktb_internal_fail(KTB_ERR_ARG, "bad device list");
    ktb_multi *m = new ktb_multi();
    m->k = k;
    if (ndev == 0) for (int d = 0; d < visible; ++d) m->devices.push_back(d);   // all of them
    else m->devices.assign(devices, devices + ndev);
    for (size_t i = 0; i < m->devices.size(); ++i)
        for (size_t j = 0; j < i; ++j)
            if (m->devices[i] == m->devices[j]) { delete m; return ktb_internal_fail(KTB_ERR_ARG, "a device is listed twice"); }
    for (int d : m->devices) {
        ktb_oligo *h = nullptr;
        if (int rc = ktb_oligo_create(k, d, &h)) {
            ktb_multi_destroy(m);
            return rc;
        }
        m->handles.push_back(h);
    }
    m->stats.assign(m->devices.size(), ktb_stats{});
    *out = m;
    return KTB_OK;
}

void ktb_multi_destroy(ktb_multi *m) {
    if (!m) return;
    for (ktb_oligo *h : m->handles) ktb_oligo_destroy(h);
    delete m;
}

int ktb_multi_device_count(const ktb_multi *m) { return m ? (int)m->devices.size() : 0; }

ktb_oligo *ktb_multi_handle(ktb_multi *m, int i) {
    return (m && i >= 0 && i < (int)m->handles.size()) ? m->handles[i] : nullptr;
}

int ktb_multi_last_stats(const ktb_multi *m, int i, ktb_stats *out, uint64_t *first_row, uint64_t *end_row) {
    if (!m || !out || i < 0 || i >= (int)m->stats.size()) return ktb_internal_fail(KTB_ERR_ARG, "bad argument");
    *out = m->stats[i];
    if (first_row) *first_row = m->bounds.size() > (size_t)i ? m->bounds[i] : 0;
    if (end_row) *end_row = m->bounds.size() > (size_t)i + 1 ? m->bounds[i + 1] : 0;
    return KTB_OK;
}

void *ktb_multi_alloc_rows(const ktb_multi *m, const uint64_t *offsets, uint64_t n, int canonical, int out_dtype) {
    if (!m || !offsets || out_dtype < 0 || out_dtype > 2) {
        ktb_internal_fail(KTB_ERR_ARG, "bad argument");
        return nullptr;
    }
    const size_t row = (size_t)ktb_oligo_dim(m->handles[0], canonical) * (out_dtype == KTB_OUT_F64 ? 8 : 4);
    const int G = (int)m->devices.size();
    std::vector<uint64_t> b(G + 1);
    if (ktb_shard_bounds(offsets, n, G, b.data())) return nullptr;
    std::vector<size_t> cuts(G + 1);
    std::vector<int> nodes(G);
    bool known = false;
    for (int g = 0; g <= G; ++g) cuts[g] = (size_t)b[g] * row;
    for (int g = 0; g < G; ++g) { nodes[g] = numa_node_of_device(m->devices[g]); known |= nodes[g] >= 0; }
    void *p = known ? alloc_mapped((size_t)n * row, cuts, nodes) : nullptr;
    return p ? p : ktb_host_alloc((size_t)n * row);
}

int ktb_multi_vectorise(ktb_multi *m, const uint8_t *bases, const uint64_t *offsets, uint64_t n, int canonical,
                        int norm_mode, int out_dtype, void *out, uint64_t *totals) {
    if (!m) return ktb_internal_fail(KTB_ERR_ARG, "null handle");
    if (n && (!offsets || !out)) return ktb_internal_fail(KTB_ERR_ARG, "null pointer");
    if (out_dtype < 0 || out_dtype > 2) return ktb_internal_fail(KTB_ERR_ARG, "unknown out_dtype");
    const int G = (int)m->devices.size();
    m->bounds.assign(G + 1, 0);
    for (auto &s : m->stats) s = ktb_stats{};
    if (n == 0) return KTB_OK;
    for (uint64_t i = 0; i < n; ++i)
        if (offsets[i + 1] < offsets[i]) return ktb_internal_fail(KTB_ERR_ARG, "offsets must be non-decreasing");
    if (int rc = ktb_shard_bounds(offsets, n, G, m->bounds.data())) return rc;
    const size_t row = (size_t)ktb_oligo_dim(m->handles[0], canonical) * (out_dtype == KTB_OUT_F64 ? 8 : 4);
    std::vector<int> rcs(G, KTB_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g] {
            const uint64_t lo = m->bounds[g], hi = m->bounds[g + 1];
            if (hi == lo) return;
            // the slice keeps the caller's absolute offsets: ktb_oligo_vectorise copies bases[offsets[lo] ..)
            rcs[g] = ktb_oligo_vectorise(m->handles[g], bases, offsets + lo, hi - lo, canonical, norm_mode, out_dtype,
                                         (uint8_t *)out + lo * row, totals ? totals + lo : nullptr);
            if (rcs[g]) errs[g] = ktb_last_error();   // thread-local: carry it to the caller's thread
            else ktb_oligo_last_stats(m->handles[g], &m->stats[g]);
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return ktb_internal_fail(rcs[g], ("device " + std::to_string(m->devices[g]) + ": " + errs[g]).c_str());
    return KTB_OK;
}

}  // extern "C"
