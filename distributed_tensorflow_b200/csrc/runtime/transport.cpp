// Native transport of the control plane (the role of TF's C++ GrpcServer / RecvTensor path).
//
// A message is a FRAME of n segments: segment 0 is the (restricted-pickle) envelope -- method name, scalars, tensor
// metadata --, segments 1..n-1 are the tensors' raw bytes.  Sending gathers all segments with writev() straight from the
// tensors' memory (no concatenation into one bytes object); receiving reads the header, lets the caller allocate one buffer
// per segment and scatters into them with readv() (tensors land in their final storage).  All calls block WITHOUT the GIL
// (ctypes releases it), wake up on a deadline so Python can honour cancellation, and never raise SIGPIPE.
//
//   header  : u32 magic 'DTF2' | u32 nseg | u64 len[nseg]          (little endian)
//   payload : segment 0 | segment 1 | ...
//
// Also here: listen / accept-with-timeout / connect-with-timeout (TCP_NODELAY on every connection: the protocol is strict
// request / reply with small envelopes) and a hang-up probe used to cancel work done on behalf of a dead client.
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <time.h>
#include <unistd.h>

#include <cstdint>
#include <vector>

namespace {

constexpr uint32_t kMagic = 0x32465444u;            // "DTF2"
constexpr int kMaxSegments = 4096;
constexpr uint64_t kMaxSegmentBytes = 1ull << 36;   // 64 GiB: a corrupt header must not drive an allocation

enum : int { DTF_NET_OK = 0, DTF_NET_EOF = -1, DTF_NET_TIMEOUT = -2, DTF_NET_BAD_FRAME = -3, DTF_NET_ERROR = -4 };

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void tune(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  setsockopt(fd, SOL_SOCKET, SO_KEEPALIVE, &one, sizeof(one));
}

// wait until fd is readable / writable; timeout_s < 0: forever.  1 ready, 0 timeout, -1 error / hang-up without data
int wait_fd(int fd, short events, double timeout_s) {
  pollfd p{fd, events, 0};
  for (;;) {
    const int ms = timeout_s < 0 ? -1 : (int)(timeout_s * 1000.0 + 0.999);
    const int r = poll(&p, 1, ms);
    if (r > 0) return (p.revents & (events | POLLHUP | POLLERR)) ? 1 : -1;
    if (r == 0) return 0;
    if (errno != EINTR) return -1;
  }
}

// read exactly n bytes; deadline < 0: none.  The deadline only applies while NOTHING of the frame has arrived yet
// (``started`` false): once a frame is under way it is read to its end (a half-read frame would desynchronise the stream).
int read_full(int fd, void* buf, size_t n, double deadline, bool* started) {
  char* p = static_cast<char*>(buf);
  size_t got = 0;
  while (got < n) {
    if (deadline >= 0 && !*started) {
      const double left = deadline - now_s();
      const int w = wait_fd(fd, POLLIN, left > 0 ? left : 0);
      if (w == 0) return DTF_NET_TIMEOUT;
      if (w < 0) return DTF_NET_EOF;
    }
    const ssize_t r = recv(fd, p + got, n - got, 0);
    if (r > 0) {
      got += (size_t)r;
      *started = true;
    } else if (r == 0) {
      return DTF_NET_EOF;
    } else if (errno == EINTR) {
      continue;
    } else if (errno == EAGAIN || errno == EWOULDBLOCK) {
      if (wait_fd(fd, POLLIN, -1) < 0) return DTF_NET_EOF;
    } else {
      return (errno == ECONNRESET || errno == EPIPE || errno == EBADF || errno == ENOTCONN) ? DTF_NET_EOF : DTF_NET_ERROR;
    }
  }
  return DTF_NET_OK;
}

}  // namespace

extern "C" {

// Listening socket bound to host:port (SO_REUSEADDR).  Returns the fd, or -errno.
int dtf_net_listen(const char* host, int port, int backlog) {
  const int fd = socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) return -errno;
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in a;
  memset(&a, 0, sizeof(a));
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, host, &a.sin_addr) != 1) {
    addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host, nullptr, &hints, &res) != 0 || res == nullptr) {
      close(fd);
      return -EADDRNOTAVAIL;
    }
    a.sin_addr = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr;
    freeaddrinfo(res);
  }
  if (bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0 || listen(fd, backlog) != 0) {
    const int e = errno;
    close(fd);
    return -e;
  }
  return fd;
}

// Next connection, or DTF_NET_TIMEOUT after timeout_s (so the accept loop can notice a shutdown), or -errno.
int dtf_net_accept(int lfd, double timeout_s) {
  const int w = wait_fd(lfd, POLLIN, timeout_s);
  if (w == 0) return DTF_NET_TIMEOUT;
  if (w < 0) return DTF_NET_ERROR;
  const int fd = accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
  if (fd < 0) return (errno == EAGAIN || errno == ECONNABORTED || errno == EINTR) ? DTF_NET_TIMEOUT : DTF_NET_ERROR;
  tune(fd);
  return fd;
}

// Connected socket, or -errno (ECONNREFUSED while the peer is not up yet: the caller retries).
int dtf_net_connect(const char* host, int port, double timeout_s) {
  sockaddr_in a;
  memset(&a, 0, sizeof(a));
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, host, &a.sin_addr) != 1) {
    addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host, nullptr, &hints, &res) != 0 || res == nullptr) return -EHOSTUNREACH;
    a.sin_addr = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr;
    freeaddrinfo(res);
  }
  const int fd = socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) return -errno;
  const int fl = fcntl(fd, F_GETFL, 0);
  fcntl(fd, F_SETFL, fl | O_NONBLOCK);
  int r = connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a));
  if (r != 0 && errno == EINPROGRESS) {
    if (wait_fd(fd, POLLOUT, timeout_s) <= 0) {
      close(fd);
      return -ETIMEDOUT;
    }
    int err = 0;
    socklen_t len = sizeof(err);
    getsockopt(fd, SOL_SOCKET, SO_ERROR, &err, &len);
    if (err != 0) {
      close(fd);
      return -err;
    }
  } else if (r != 0) {
    const int e = errno;
    close(fd);
    return -e;
  }
  fcntl(fd, F_SETFL, fl);
  tune(fd);
  return fd;
}

// One frame: header + every segment, gathered with writev (MSG_NOSIGNAL semantics via sendmsg).
int dtf_net_send(int fd, const void* const* bufs, const uint64_t* lens, int n) {
  if (n <= 0 || n > kMaxSegments) return DTF_NET_BAD_FRAME;
  std::vector<uint64_t> head(1 + (size_t)n);
  head[0] = (uint64_t)kMagic | ((uint64_t)(uint32_t)n << 32);
  for (int i = 0; i < n; ++i) head[1 + i] = lens[i];
  std::vector<iovec> iov;
  iov.reserve((size_t)n + 1);
  iov.push_back({head.data(), head.size() * sizeof(uint64_t)});
  for (int i = 0; i < n; ++i)
    if (lens[i]) iov.push_back({const_cast<void*>(bufs[i]), (size_t)lens[i]});
  size_t at = 0;
  while (at < iov.size()) {
    msghdr m;
    memset(&m, 0, sizeof(m));
    m.msg_iov = &iov[at];
    m.msg_iovlen = iov.size() - at > 512 ? 512 : iov.size() - at;
    ssize_t w = sendmsg(fd, &m, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) {
        if (wait_fd(fd, POLLOUT, -1) < 0) return DTF_NET_EOF;
        continue;
      }
      return (errno == EPIPE || errno == ECONNRESET || errno == EBADF || errno == ENOTCONN) ? DTF_NET_EOF : DTF_NET_ERROR;
    }
    size_t left = (size_t)w;
    while (left > 0 && at < iov.size()) {
      if (left >= iov[at].iov_len) {
        left -= iov[at].iov_len;
        ++at;
      } else {
        iov[at].iov_base = static_cast<char*>(iov[at].iov_base) + left;
        iov[at].iov_len -= left;
        left = 0;
      }
    }
  }
  return DTF_NET_OK;
}

// Header of the next frame: fills lens[0..n) and returns n (> 0); DTF_NET_TIMEOUT when nothing arrived within timeout_s
// (timeout_s < 0: wait forever), DTF_NET_EOF when the peer closed, DTF_NET_BAD_FRAME for a stream that is not ours.
int dtf_net_recv_header(int fd, uint64_t* lens, int max_n, double timeout_s) {
  uint64_t first = 0;
  bool started = false;
  const double deadline = timeout_s < 0 ? -1.0 : now_s() + timeout_s;
  int rc = read_full(fd, &first, sizeof(first), deadline, &started);
  if (rc != DTF_NET_OK) return rc;
  if ((uint32_t)(first & 0xFFFFFFFFu) != kMagic) return DTF_NET_BAD_FRAME;
  const int n = (int)(first >> 32);
  if (n <= 0 || n > kMaxSegments || n > max_n) return DTF_NET_BAD_FRAME;
  rc = read_full(fd, lens, sizeof(uint64_t) * (size_t)n, -1.0, &started);
  if (rc != DTF_NET_OK) return rc == DTF_NET_TIMEOUT ? DTF_NET_EOF : rc;
  for (int i = 0; i < n; ++i)
    if (lens[i] > kMaxSegmentBytes) return DTF_NET_BAD_FRAME;
  return n;
}

// Body of the frame whose header was just read: scatter into the caller's buffers.
int dtf_net_recv_body(int fd, void* const* bufs, const uint64_t* lens, int n) {
  bool started = true;
  for (int i = 0; i < n; ++i) {
    if (!lens[i]) continue;
    const int rc = read_full(fd, bufs[i], (size_t)lens[i], -1.0, &started);
    if (rc != DTF_NET_OK) return rc == DTF_NET_TIMEOUT ? DTF_NET_EOF : rc;
  }
  return DTF_NET_OK;
}

// 1 once the peer has hung up (process killed, socket closed) -- also while a handler is busy with its request.
int dtf_net_peer_closed(int fd) {
  pollfd p{fd, (short)(POLLRDHUP | POLLHUP | POLLERR), 0};
  const int r = poll(&p, 1, 0);
  if (r < 0) return errno == EINTR ? 0 : 1;
  return (r > 0 && (p.revents & (POLLRDHUP | POLLHUP | POLLERR | POLLNVAL))) ? 1 : 0;
}

int dtf_net_local_port(int fd) {
  sockaddr_in a;
  socklen_t len = sizeof(a);
  if (getsockname(fd, reinterpret_cast<sockaddr*>(&a), &len) != 0) return -errno;
  return (int)ntohs(a.sin_port);
}

void dtf_net_shutdown(int fd) { shutdown(fd, SHUT_RDWR); }
void dtf_net_close(int fd) { close(fd); }

}  // extern "C"
