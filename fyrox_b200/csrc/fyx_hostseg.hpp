// fyx_hostseg.hpp — one host memory segment shared by the ranks of a node (host-only code, no CUDA).
//
// Multi-GPU frames end with every rank holding the all-gathered visible lists on its device; the HOST copy
// of the whole lists is assembled here instead of being pulled by one rank over one PCIe link: rank r
// copies ITS OWN compacted list of frustum f to  lists(slot, f) + offset[f][r]  of a segment every rank has
// mapped (and registered with CUDA, so the copies are plain DMA), offset[f][r] = sum of the counts of the
// ranks before it.  N PCIe links work in parallel and each entry crosses PCIe exactly once per node.
// The reference has no counterpart (single process); its consumer of the list is the CPU loop of
// RenderDataBundleStorage::from_graph (renderer/bundle.rs:988-1004).
//
// Backing: memfd_create (not limited by the size of /dev/shm); the creator's (pid, fd) travel to the other
// ranks, which open /proc/<pid>/fd/<fd>.  Frames alternate between two slots (epoch & 1).
// Protocol per gathered frame with epoch e (every rank runs the same sequence of gathered frames):
//   begin(e)        rank r declares it no longer reads epoch e-2 (same slot): released[slot][r] = e
//   wait_writable(e) before writing: all ranks have begun e
//   ... rank r DMA-copies its own lists into the slot ...
//   publish(e)      done[slot][r] = e  (release)
//   wait_complete(e) before reading the whole lists: done[slot][q] >= e for every q (acquire)
#pragma once
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace fyx {

constexpr int kSegMaxRanks = 16;
constexpr size_t kSegHeaderBytes = 4096;

struct HostSegHeader {
    std::atomic<uint64_t> magic;
    std::atomic<uint64_t> done[2][kSegMaxRanks];
    std::atomic<uint64_t> released[2][kSegMaxRanks];
};
static_assert(sizeof(HostSegHeader) <= kSegHeaderBytes, "header fits its page");
static_assert(std::atomic<uint64_t>::is_always_lock_free, "flags are plain 64-bit words");

class HostSeg {
public:
    int fd = -1;
    bool owner = false;
    unsigned char *base = nullptr;
    size_t bytes = 0;
    uint64_t cap_total = 0; // entries per (slot, frustum) region
    uint32_t nf_cap = 0;
    int nranks = 1, rank = 0;
    std::string err;

    static size_t bytes_for(uint64_t cap_total, uint32_t nf_cap) { return kSegHeaderBytes + (size_t)2 * nf_cap * cap_total * sizeof(uint32_t); }

    // creator side: returns (pid, fd) through out_pid_fd
    bool create(uint64_t cap, uint32_t nf, int nranks_, int rank_, int64_t out_pid_fd[2])
    {
        close();
        cap_total = cap; nf_cap = nf; nranks = nranks_; rank = rank_;
        bytes = bytes_for(cap, nf);
        fd = (int)syscall(SYS_memfd_create, "fyx_hostseg", 0u);
        if (fd < 0) return fail("memfd_create");
        if (ftruncate(fd, (off_t)bytes) != 0) return fail("ftruncate");
        if (!map()) return false;
        owner = true;
        interleave_pages(); // before the first touch
        new (header()) HostSegHeader();
        for (int s = 0; s < 2; ++s)
            for (int r = 0; r < kSegMaxRanks; ++r) {
                header()->done[s][r].store(0, std::memory_order_relaxed);
                header()->released[s][r].store(0, std::memory_order_relaxed);
            }
        header()->magic.store(0x46595853454731ull, std::memory_order_release);
        out_pid_fd[0] = (int64_t)getpid();
        out_pid_fd[1] = fd;
        return true;
    }

    // other ranks: open the creator's descriptor through /proc
    bool open_from(const int64_t pid_fd[2], uint64_t cap, uint32_t nf, int nranks_, int rank_)
    {
        close();
        cap_total = cap; nf_cap = nf; nranks = nranks_; rank = rank_;
        bytes = bytes_for(cap, nf);
        char path[64];
        snprintf(path, sizeof path, "/proc/%lld/fd/%lld", (long long)pid_fd[0], (long long)pid_fd[1]);
        fd = ::open(path, O_RDWR);
        if (fd < 0) return fail(path);
        if (!map()) return false;
        if (header()->magic.load(std::memory_order_acquire) != 0x46595853454731ull) {
            err = "host segment: bad magic";
            close();
            return false;
        }
        return true;
    }

    void close()
    {
        if (base) munmap(base, bytes);
        base = nullptr;
        if (fd >= 0) ::close(fd);
        fd = -1;
        owner = false;
    }
    ~HostSeg() { close(); }

    HostSegHeader *header() const { return reinterpret_cast<HostSegHeader *>(base); }
    uint32_t *list(uint64_t epoch, uint32_t f) const
    {
        return reinterpret_cast<uint32_t *>(base + kSegHeaderBytes) + ((size_t)(epoch & 1) * nf_cap + f) * cap_total;
    }

    void begin(uint64_t epoch) { header()->released[epoch & 1][rank].store(epoch, std::memory_order_release); }
    void publish(uint64_t epoch) { header()->done[epoch & 1][rank].store(epoch, std::memory_order_release); }
    bool wait_writable(uint64_t epoch, double timeout_s = 20.0) { return wait_all(header()->released[epoch & 1], epoch, timeout_s, "released"); }
    bool wait_complete(uint64_t epoch, double timeout_s = 20.0) { return wait_all(header()->done[epoch & 1], epoch, timeout_s, "done"); }

private:
    // The segment is written by GPUs on every socket of the node: spread its pages over all NUMA nodes instead of letting
    // the creator's first touch (cudaHostRegister pins them) put them all next to rank 0.  Best effort (mbind may be denied).
    void interleave_pages()
    {
#ifdef SYS_mbind
        unsigned long mask[16];
        int nodes = 0;
        for (int n = 0; n < 64; ++n) {
            char path[64];
            snprintf(path, sizeof path, "/sys/devices/system/node/node%d", n);
            if (access(path, F_OK) == 0) nodes = n + 1;
        }
        if (nodes < 2) return;
        memset(mask, 0, sizeof mask);
        for (int n = 0; n < nodes; ++n) mask[n / (8 * sizeof(unsigned long))] |= 1ul << (n % (8 * sizeof(unsigned long)));
        const int MPOL_INTERLEAVE_ = 3;
        (void)syscall(SYS_mbind, base, bytes, MPOL_INTERLEAVE_, mask, (unsigned long)(nodes + 1), 0u);
#endif
    }
    bool map()
    {
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (p == MAP_FAILED) return fail("mmap");
        base = static_cast<unsigned char *>(p);
        return true;
    }
    bool fail(const char *what)
    {
        err = std::string("host segment: ") + what + ": " + strerror(errno);
        close();
        return false;
    }
    bool wait_all(std::atomic<uint64_t> *flags, uint64_t epoch, double timeout_s, const char *what)
    {
        const auto t0 = std::chrono::steady_clock::now();
        for (int q = 0; q < nranks; ++q) {
            int spins = 0;
            while (flags[q].load(std::memory_order_acquire) < epoch) {
                if (++spins > 200) {
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
                    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
                        err = std::string("host segment: timed out waiting for rank ") + std::to_string(q) + " (" + what + ", epoch " + std::to_string(epoch) + ")";
                        return false;
                    }
                }
            }
        }
        return true;
    }
};

} // namespace fyx
