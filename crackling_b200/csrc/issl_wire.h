// issl_wire.h -- the request/response framing between bin/isslScoreOfftargets (client mode) and
// bin/isslScoreServer, the resident scorer that keeps the index in HBM between invocations.
//
// Why it exists: Crackling's pipeline starts the scorer once per page of guides
// (/root/reference/src/crackling/Crackling.py:737-778) and the reference re-reads the whole .issl
// every time (isslScoreOfftargets.cpp:152-243, ~28 GB at human scale).  With scoring at ~0.1 s per
// 100 000 guides, that reload is the whole cost; the server pays it once.
//
// Transport: a unix stream socket (path in ISSL_SERVER).  One request per connection.  All fields
// little-endian, fixed-width.  Host programs only -- nothing here is part of the C ABI.
#ifndef ISSL_WIRE_H
#define ISSL_WIRE_H

#include <cerrno>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/un.h>
#include <unistd.h>

namespace issl_wire {

constexpr char kReqMagic[8] = {'I', 'S', 'S', 'L', 'R', 'E', 'Q', '1'};
constexpr char kRspMagic[8] = {'I', 'S', 'S', 'L', 'R', 'S', 'P', '1'};
constexpr int kMaxDevices = 16;

enum Op : uint32_t { kScore = 1, kPing = 2, kShutdown = 3, kDrop = 4 /* forget every cached index */ };

struct Request {
    char magic[8];
    uint32_t op;
    int32_t maxDist;
    double threshold;
    int32_t method;
    int32_t layout;
    uint64_t nGuides;
    uint32_t pathLen;        // bytes of the index path that follow (no terminator)
    uint32_t nDevices;       // 0 = server's choice
    int32_t devices[kMaxDevices];
    // then: path[pathLen], guides[nGuides] (u64 each)
};

struct Response {
    char magic[8];
    int32_t status;          // issl_status
    uint32_t msgLen;         // bytes of diagnostic text that follow
    uint64_t n;              // scores per column that follow the text (0 on error)
    double loadSeconds;      // time spent bringing the index into HBM for this request (0 when cached)
    double scoreSeconds;
    uint32_t cached;         // 1 when the index was already resident
    uint32_t nDevices;
    uint64_t candidates, hits, earlyExits;
    // then: msg[msgLen], mit[n] (f64), cfd[n] (f64)
};

inline bool read_full(int fd, void *buf, size_t n)
{
    char *p = static_cast<char *>(buf);
    while (n) {
        const ssize_t r = ::read(fd, p, n);
        if (r == 0) return false;
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}

inline bool write_full(int fd, const void *buf, size_t n)
{
    const char *p = static_cast<const char *>(buf);
    while (n) {
        const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}

inline bool fill_addr(const char *path, sockaddr_un *addr)
{
    memset(addr, 0, sizeof *addr);
    addr->sun_family = AF_UNIX;
    if (strlen(path) >= sizeof addr->sun_path) return false;
    strcpy(addr->sun_path, path);
    return true;
}

// connected socket or -1
inline int connect_to(const char *path)
{
    sockaddr_un addr;
    if (!fill_addr(path, &addr)) return -1;
    const int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return -1;
    if (connect(fd, reinterpret_cast<sockaddr *>(&addr), sizeof addr) != 0) { close(fd); return -1; }
    return fd;
}

}  // namespace issl_wire
#endif
