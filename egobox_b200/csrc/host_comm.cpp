// The one exchange of the sharded path, behind the C ABI: after every rank has fitted its share of the multistart
// chains / theta candidates / experts, the ranks exchange (value, payload[h]) -- 8 (h + 2) bytes each -- and keep the best
// (the reduction of gp/src/algorithm.rs:942-945 across processes; SURVEY.md section 8 (b) / (e): egx_comm_init,
// egx_argmin_allreduce).  The data path itself has no collective, and a message of a few dozen bytes is latency, not
// bandwidth: the exchange runs over a TCP star (rank 0 listens, the torchrun MASTER_ADDR convention) so that a Rust / C
// caller needs neither torch.distributed nor its own NCCL binding.  Callers that already hold a torch process group use
// egobox_b200/parallel.py (NCCL all_gather) instead; both produce the same winner.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

struct egx_comm {
    int nranks = 1, rank = 0;
    int fd = -1;                 // rank > 0: connection to rank 0
    std::vector<int> peers;      // rank 0: connection of every rank (index = rank, [0] unused)
    int listen_fd = -1;
};

namespace {

bool send_all(int fd, const void* buf, size_t len) {
    const char* p = static_cast<const char*>(buf);
    while (len > 0) {
        const ssize_t k = ::send(fd, p, len, MSG_NOSIGNAL);
        if (k <= 0) return false;
        p += k;
        len -= static_cast<size_t>(k);
    }
    return true;
}
bool recv_all(int fd, void* buf, size_t len) {
    char* p = static_cast<char*>(buf);
    while (len > 0) {
        const ssize_t k = ::recv(fd, p, len, 0);
        if (k <= 0) return false;
        p += k;
        len -= static_cast<size_t>(k);
    }
    return true;
}
void set_timeouts(int fd, int timeout_ms) {
    timeval tv;
    tv.tv_sec = timeout_ms / 1000;
    tv.tv_usec = (timeout_ms % 1000) * 1000;
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}
void close_all(egx_comm* c) {
    if (c->fd >= 0) ::close(c->fd);
    for (int fd : c->peers)
        if (fd >= 0) ::close(fd);
    if (c->listen_fd >= 0) ::close(c->listen_fd);
    c->fd = c->listen_fd = -1;
    c->peers.clear();
}
bool resolve(const char* addr, int port, sockaddr_in* out) {
    std::memset(out, 0, sizeof(*out));
    out->sin_family = AF_INET;
    out->sin_port = htons(static_cast<uint16_t>(port));
    if (inet_pton(AF_INET, addr, &out->sin_addr) == 1) return true;
    addrinfo hints, *res = nullptr;
    std::memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(addr, nullptr, &hints, &res) != 0 || res == nullptr) return false;
    out->sin_addr = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr;
    freeaddrinfo(res);
    return true;
}

}  // namespace

extern "C" int egx_comm_init(egx_comm** out, int nranks, int rank, const char* addr, int port, int timeout_ms) try {
    if (!out) return EGX_INVALID_VALUE;
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && (!addr || port < 1 || port > 65535))) {
        egx_set_error("egx_comm_init: bad arguments (nranks=%d rank=%d port=%d)", nranks, rank, port);
        return EGX_INVALID_VALUE;
    }
    if (timeout_ms <= 0) timeout_ms = 60000;
    egx_comm* c = new egx_comm();
    c->nranks = nranks;
    c->rank = rank;
    if (nranks == 1) {
        *out = c;
        return EGX_OK;
    }
    sockaddr_in sa;
    if (!resolve(addr, port, &sa)) {
        egx_set_error("egx_comm_init: cannot resolve %s", addr);
        delete c;
        return EGX_INVALID_VALUE;
    }
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    if (rank == 0) {
        c->peers.assign(nranks, -1);
        c->listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
        int one = 1;
        setsockopt(c->listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        sockaddr_in any = sa;
        any.sin_addr.s_addr = htonl(INADDR_ANY);
        // rank 0's own address by convention: listen there (loopback stays loopback); any interface if it is not local
        if (c->listen_fd < 0 ||
            (::bind(c->listen_fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0 &&
             ::bind(c->listen_fd, reinterpret_cast<sockaddr*>(&any), sizeof(any)) != 0) ||
            ::listen(c->listen_fd, nranks) != 0) {
            egx_set_error("egx_comm_init: rank 0 cannot listen on port %d (%s)", port, std::strerror(errno));
            close_all(c);
            delete c;
            return EGX_CUDA_ERROR;
        }
        timeval tv;
        tv.tv_sec = timeout_ms / 1000;
        tv.tv_usec = (timeout_ms % 1000) * 1000;
        setsockopt(c->listen_fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));      // bounds accept()
        for (int k = 1; k < nranks; ++k) {
            const int fd = ::accept(c->listen_fd, nullptr, nullptr);
            int32_t peer = -1;
            if (fd >= 0) set_timeouts(fd, timeout_ms);
            if (fd < 0 || !recv_all(fd, &peer, sizeof(peer)) || peer < 1 || peer >= nranks || c->peers[peer] >= 0) {
                egx_set_error("egx_comm_init: rank 0 saw %d of %d peers before the timeout (or a bad hello)", k - 1, nranks - 1);
                if (fd >= 0) ::close(fd);
                close_all(c);
                delete c;
                return EGX_CUDA_ERROR;
            }
            c->peers[peer] = fd;
        }
        ::close(c->listen_fd);
        c->listen_fd = -1;
    } else {
        for (;;) {
            c->fd = ::socket(AF_INET, SOCK_STREAM, 0);
            if (c->fd >= 0 && ::connect(c->fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) == 0) break;
            if (c->fd >= 0) ::close(c->fd);
            c->fd = -1;
            if (std::chrono::steady_clock::now() > deadline) {
                egx_set_error("egx_comm_init: rank %d cannot reach rank 0 at %s:%d", rank, addr, port);
                delete c;
                return EGX_CUDA_ERROR;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(20));          // rank 0 may not be listening yet
        }
        set_timeouts(c->fd, timeout_ms);
        const int32_t me = rank;
        if (!send_all(c->fd, &me, sizeof(me))) {
            egx_set_error("egx_comm_init: hello to rank 0 failed");
            close_all(c);
            delete c;
            return EGX_CUDA_ERROR;
        }
    }
    *out = c;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" void egx_comm_destroy(egx_comm* c) {
    if (!c) return;
    close_all(c);
    delete c;
}

extern "C" int egx_comm_rank(const egx_comm* c) { return c ? c->rank : -1; }
extern "C" int egx_comm_size(const egx_comm* c) { return c ? c->nranks : 0; }

// every rank contributes `count` doubles, every rank receives nranks * count doubles in rank order
extern "C" int egx_comm_allgather(egx_comm* c, const double* send, int count, double* recv) try {
    if (!c || !send || !recv || count < 1) return EGX_INVALID_VALUE;
    const size_t bytes = static_cast<size_t>(count) * sizeof(double);
    if (c->nranks == 1) {
        std::memcpy(recv, send, bytes);
        return EGX_OK;
    }
    bool ok = true;
    if (c->rank == 0) {
        std::memcpy(recv, send, bytes);
        for (int k = 1; k < c->nranks && ok; ++k) ok = recv_all(c->peers[k], recv + static_cast<size_t>(k) * count, bytes);
        for (int k = 1; k < c->nranks && ok; ++k) ok = send_all(c->peers[k], recv, bytes * c->nranks);
    } else {
        ok = send_all(c->fd, send, bytes) && recv_all(c->fd, recv, bytes * c->nranks);
    }
    if (!ok) {
        egx_set_error("egx_comm_allgather: a peer went away or timed out (rank %d of %d)", c->rank, c->nranks);
        return EGX_CUDA_ERROR;
    }
    return EGX_OK;
}
EGX_ABI_CATCH

// In place: (value, payload[h]) of the rank with the smallest value (first strictly smaller wins in rank order, exactly
// the fold of algorithm.rs:942-945; NaN never wins; when no rank has a finite-or-infinite comparable value rank 0's pair
// stays).  winner_rank may be NULL.
extern "C" int egx_argmin_allreduce(egx_comm* c, double* value, double* payload, int h, int* winner_rank) try {
    if (!c || !value || h < 0 || (h > 0 && !payload)) return EGX_INVALID_VALUE;
    const int w = h + 1;
    std::vector<double> mine(w), all(static_cast<size_t>(w) * c->nranks);
    mine[0] = *value;
    for (int i = 0; i < h; ++i) mine[1 + i] = payload[i];
    const int st = egx_comm_allgather(c, mine.data(), w, all.data());
    if (st != EGX_OK) return st;
    int best = 0;
    for (int k = 1; k < c->nranks; ++k) {
        const double v = all[static_cast<size_t>(k) * w], b = all[static_cast<size_t>(best) * w];
        if (v < b || (std::isnan(b) && !std::isnan(v))) best = k;
    }
    *value = all[static_cast<size_t>(best) * w];
    for (int i = 0; i < h; ++i) payload[i] = all[static_cast<size_t>(best) * w + 1 + i];
    if (winner_rank) *winner_rank = best;
    return EGX_OK;
}
EGX_ABI_CATCH
