// Host-side lab for the drop-in's last copy: page cache -> staging, 12.2 MB per frame, T threads.
//   A  pread straight into the staging slot (what lfd_fits_load_frame did in round 1 / early round 2)
//   B  mmap + memcpy
//   C  mmap + non-temporal (streaming) stores
//   E  as C with 16-byte SSE2 streaming stores (what the library uses: no -mavx2 in the build)
//   D  pread into a small cache-resident bounce buffer, then non-temporal stores into the slot
// The kernel's copy_to_user writes with ordinary stores: every destination line is read for ownership before it is
// written, so A moves 3 bytes over the memory bus per payload byte (read source, read destination, write destination);
// C and D move 2.  build: g++ -O2 -mavx2 -pthread -o ingest_lab ingest_lab.cpp ; run: ./ingest_lab <dir> <nfiles> <threads> <variant> [reps]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <immintrin.h>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>

static const size_t PAYLOAD = (size_t)1489 * 2048 * 4;
static const size_t HDR = 2880 * 3;

static void stream_copy(char* dst, const char* src, size_t n)
{
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 31); i++) dst[i] = src[i];
    for (; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
}

static void stream_copy_sse2(char* dst, const char* src, size_t n)      // baseline x86-64: no -mavx2 needed
{
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 15); i++) dst[i] = src[i];
    for (; i + 64 <= n; i += 64) {
        __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a); _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c); _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
}

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s dir nfiles threads variant [reps] [bounce_kb]\n", argv[0]); return 2; }
    std::string dir = argv[1];
    int nfiles = atoi(argv[2]), T = atoi(argv[3]);
    char variant = argv[4][0];
    int reps = argc > 5 ? atoi(argv[5]) : 4;
    size_t bounce = (argc > 6 ? atoi(argv[6]) : 256) * 1024;
    // files: created once (random-ish content), then read once so they sit in the page cache
    std::vector<std::string> paths;
    for (int i = 0; i < nfiles; i++) {
        std::string p = dir + "/f" + std::to_string(i) + ".bin";
        paths.push_back(p);
        if (access(p.c_str(), R_OK) != 0) {
            std::vector<char> buf(HDR + PAYLOAD);
            for (size_t k = 0; k < buf.size(); k += 8) *(uint64_t*)&buf[k] = k * 0x9E3779B97F4A7C15ull + i;
            int fd = open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fd < 0 || write(fd, buf.data(), buf.size()) != (ssize_t)buf.size()) { perror("write"); return 1; }
            close(fd);
        }
    }
    const int slots = 96;                                  // three batches of 32 frames, like the handle ring
    char* staging = (char*)aligned_alloc(4096, (size_t)slots * PAYLOAD);
    memset(staging, 1, (size_t)slots * PAYLOAD);
    const int frames = nfiles * reps;
    std::atomic<int> next{0};
    std::atomic<unsigned long long> sum{0};
    auto worker = [&]() {
        char* bb = (char*)aligned_alloc(4096, bounce);
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= frames) break;
            char* dst = staging + (size_t)(i % slots) * PAYLOAD;
            int fd = open(paths[i % nfiles].c_str(), O_RDONLY);
            if (fd < 0) { perror("open"); exit(1); }
            if (variant == 'A') {
                size_t got = 0;
                while (got < PAYLOAD) { ssize_t k = pread(fd, dst + got, PAYLOAD - got, HDR + got); if (k <= 0) { perror("pread"); exit(1); } got += k; }
            } else if (variant == 'B' || variant == 'C' || variant == 'E') {
                char* m = (char*)mmap(nullptr, HDR + PAYLOAD, PROT_READ, MAP_SHARED | MAP_POPULATE, fd, 0);
                if (m == MAP_FAILED) { perror("mmap"); exit(1); }
                if (variant == 'B') memcpy(dst, m + HDR, PAYLOAD); else if (variant == 'C') stream_copy(dst, m + HDR, PAYLOAD); else stream_copy_sse2(dst, m + HDR, PAYLOAD);
                munmap(m, HDR + PAYLOAD);
            } else {
                size_t got = 0;
                while (got < PAYLOAD) {
                    size_t want = PAYLOAD - got < bounce ? PAYLOAD - got : bounce;
                    ssize_t k = pread(fd, bb, want, HDR + got);
                    if (k <= 0) { perror("pread"); exit(1); }
                    stream_copy(dst + got, bb, (size_t)k);
                    got += k;
                }
            }
            close(fd);
            sum += (unsigned char)dst[12345];
        }
        free(bb);
    };
    for (int pass = 0; pass < 2; pass++) {                 // pass 0 warms the page cache
        next = 0;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(worker);
        for (auto& t : th) t.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (pass) printf("variant %c threads %2d bounce %4zu KB: %7.0f frames/s  %6.1f GB/s payload  (check %llu)\n", variant, T, bounce / 1024, frames / s, frames * (double)PAYLOAD / s / 1e9, sum.load());
    }
    return 0;
}
