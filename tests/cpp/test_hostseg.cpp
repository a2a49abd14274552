// test_hostseg.cpp — CPU-only test of the node-wide host segment the multi-GPU frames assemble their visible lists in
// (fyrox_b200/csrc/fyx_hostseg.hpp): N processes (fork) act as the ranks; every epoch each writes its own part of every
// frustum's list at its offset, publishes, waits for the others and checks the WHOLE lists — while already racing ahead
// into the next epochs as far as the two-slot protocol allows.  A torn or overwritten list fails the check.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <sys/wait.h>

#include "../../fyrox_b200/csrc/fyx_hostseg.hpp"

using namespace fyx;

static uint32_t value_of(int rank, uint64_t epoch, uint32_t f, uint32_t i) { return (uint32_t)(rank * 1000003u + epoch * 7919u + f * 104729u + i); }
static uint32_t count_of(int rank, uint64_t epoch, uint32_t f) { return (uint32_t)((rank * 37u + epoch * 11u + f * 5u) % 900u); }

static int run_rank(HostSeg &seg, int nranks, int rank, uint32_t nf, int epochs)
{
    std::mt19937 rng(rank * 17 + 1);
    for (uint64_t e = 1; e <= (uint64_t)epochs; ++e) {
        seg.begin(e);
        if (rng() % 3 == 0) usleep(rng() % 300); // ranks drift against each other
        if (!seg.wait_writable(e, 10.0)) { fprintf(stderr, "rank %d: %s\n", rank, seg.err.c_str()); return 2; }
        for (uint32_t f = 0; f < nf; ++f) {
            uint32_t off = 0;
            for (int q = 0; q < rank; ++q) off += count_of(q, e, f);
            uint32_t *dst = seg.list(e, f) + off;
            const uint32_t n = count_of(rank, e, f);
            for (uint32_t i = 0; i < n; ++i) dst[i] = value_of(rank, e, f, i);
        }
        seg.publish(e);
        if (rank == 0 || e % 3 == 0) { // the consumer (and sometimes everybody) reads the whole lists
            if (!seg.wait_complete(e, 10.0)) { fprintf(stderr, "rank %d: %s\n", rank, seg.err.c_str()); return 3; }
            for (uint32_t f = 0; f < nf; ++f) {
                const uint32_t *src = seg.list(e, f);
                uint32_t off = 0;
                for (int q = 0; q < nranks; ++q) {
                    const uint32_t n = count_of(q, e, f);
                    for (uint32_t i = 0; i < n; ++i)
                        if (src[off + i] != value_of(q, e, f, i)) {
                            fprintf(stderr, "rank %d: epoch %llu frustum %u: entry %u of rank %d is wrong\n", rank, (unsigned long long)e, f, i, q);
                            return 4;
                        }
                    off += n;
                }
            }
        }
    }
    return 0;
}

int main()
{
    const int nranks = 4, epochs = 400;
    const uint32_t nf = 6;
    HostSeg seg;
    int64_t pid_fd[2];
    if (!seg.create(900 * nranks, nf, nranks, 0, pid_fd)) { fprintf(stderr, "create: %s\n", seg.err.c_str()); return 1; }
    std::vector<pid_t> kids;
    for (int r = 1; r < nranks; ++r) {
        const pid_t p = fork();
        if (p == 0) {
            HostSeg mine; // a fresh mapping through /proc/<pid>/fd/<fd>, as a separately launched rank would make it
            if (!mine.open_from(pid_fd, 900 * nranks, nf, nranks, r)) { fprintf(stderr, "open_from: %s\n", mine.err.c_str()); _exit(9); }
            _exit(run_rank(mine, nranks, r, nf, epochs));
        }
        kids.push_back(p);
    }
    int rc = run_rank(seg, nranks, 0, nf, epochs);
    for (pid_t p : kids) {
        int st = 0;
        waitpid(p, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 100 + (WIFEXITED(st) ? WEXITSTATUS(st) : 99);
    }
    // a second segment size / a wrong descriptor are reported, not fatal
    HostSeg bad;
    int64_t nowhere[2] = {1, 987654};
    if (bad.open_from(nowhere, 16, 1, 2, 1)) { fprintf(stderr, "open_from of a bogus descriptor succeeded\n"); rc = rc ? rc : 50; }
    if (rc == 0) printf("hostseg ok: %d ranks x %d epochs x %u frusta\n", nranks, epochs, nf);
    return rc;
}
