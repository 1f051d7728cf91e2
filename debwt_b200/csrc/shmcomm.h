// Host-side plumbing of the sharded build (shard.cu): a communicator for ranks on ONE node -- separate processes (one per
// GPU, the way torchrun starts them) or threads of one process (host/deBWT -g 0,1,...) -- over a POSIX shared-memory
// segment.  Only small control data goes through it (counts, IPC handles, splitter samples); every byte of the data path
// moves GPU to GPU over NVLink.  The reference has no distributed code at all.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <string>

namespace debwt {

class ShmComm {
public:
    static constexpr size_t kSlot = 64 << 10;          // bytes one rank can contribute to one allgather
    int rank = 0, world = 1;

    // Collective.  Rank 0 creates the segment, the others attach; `tag` must be unique to this group of ranks.
    int open(const std::string& tag, int rank_, int world_, std::string* err) {
        rank = rank_; world = world_;
        if (world == 1) return 0;
        name_ = "/debwt_" + tag;
        bytes_ = sizeof(Header) + (size_t)world * kSlot;
        int fd = -1;
        if (rank == 0) {
            shm_unlink(name_.c_str());
            fd = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)bytes_) != 0) { *err = "shm_open/ftruncate failed for " + name_; return -1; }
        } else {
            const double t0 = now();
            for (;;) {
                fd = shm_open(name_.c_str(), O_RDWR, 0600);
                struct stat sb;
                if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size >= bytes_) break;
                if (fd >= 0) close(fd);
                if (now() - t0 > 120.0) { *err = "timed out waiting for rank 0 to create " + name_; return -1; }
                usleep(1000);
            }
        }
        void* p = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) { *err = "mmap failed for " + name_; return -1; }
        hdr_ = static_cast<Header*>(p);
        slots_ = reinterpret_cast<unsigned char*>(p) + sizeof(Header);
        if (rank == 0) {
            hdr_->arrived = 0; hdr_->generation = 0; hdr_->world = (uint32_t)world;
            __atomic_store_n(&hdr_->magic, kMagic, __ATOMIC_RELEASE);
        } else {
            const double t0 = now();
            while (__atomic_load_n(&hdr_->magic, __ATOMIC_ACQUIRE) != kMagic) {
                if (now() - t0 > 120.0) { *err = "timed out waiting for rank 0 to initialise " + name_; return -1; }
                usleep(200);
            }
        }
        if (barrier(err)) return -1;
        if (rank == 0) shm_unlink(name_.c_str());      // everybody is attached: the name is not needed any more
        return 0;
    }

    void close_all() {
        if (hdr_) munmap(hdr_, bytes_);
        hdr_ = nullptr;
    }

    int barrier(std::string* err) {
        if (world == 1) return 0;
        const uint32_t g = __atomic_load_n(&hdr_->generation, __ATOMIC_ACQUIRE);
        if (__atomic_add_fetch(&hdr_->arrived, 1u, __ATOMIC_ACQ_REL) == (uint32_t)world) {
            __atomic_store_n(&hdr_->arrived, 0u, __ATOMIC_RELAXED);
            __atomic_store_n(&hdr_->generation, g + 1, __ATOMIC_RELEASE);
            return 0;
        }
        const double t0 = now();
        unsigned spins = 0;
        while (__atomic_load_n(&hdr_->generation, __ATOMIC_ACQUIRE) == g) {
            if (++spins > 2000) {
                sched_yield();
                if ((spins & 1023) == 0 && now() - t0 > 600.0) { if (err) *err = "barrier timed out (a rank died?)"; return -1; }
            }
        }
        return 0;
    }

    // all[r * bytes .. ) = rank r's `mine`; bytes <= kSlot
    int allgather(const void* mine, size_t bytes, void* all, std::string* err) {
        if (world == 1) { memcpy(all, mine, bytes); return 0; }
        if (bytes > kSlot) { if (err) *err = "allgather: contribution larger than a slot"; return -1; }
        memcpy(slots_ + (size_t)rank * kSlot, mine, bytes);
        if (barrier(err)) return -1;
        for (int r = 0; r < world; ++r) memcpy(static_cast<unsigned char*>(all) + (size_t)r * bytes, slots_ + (size_t)r * kSlot, bytes);
        return barrier(err);
    }

private:
    static constexpr uint32_t kMagic = 0xdeb70b20u;
    struct Header {
        uint32_t magic, world;
        uint32_t arrived, generation;
        unsigned char pad[48];
    };
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    }
    std::string name_;
    size_t bytes_ = 0;
    Header* hdr_ = nullptr;
    unsigned char* slots_ = nullptr;
};

}  // namespace debwt
