// Sequence numbers of the peer-memory interface exchange (host-side bookkeeping, see api.cu / kernels.cuh).
//
// Exchange number s uses receive slot s & 1; consecutive REAL exchanges must alternate slots (a rank can
// run at most one exchange ahead of a peer).  Launches that the device skips after convergence consume
// numbers without exchanging, so after a stretch of `enqueued` iterations of which only `run` really
// executed the next number is re-aligned: it must differ in parity from the last real exchange.
// Pure function (every rank evaluates it on identical arguments): returns the counter value to continue from.
#pragma once
#include <algorithm>
#include <cstdint>

namespace kb {

inline uint32_t realigned_exchange_seq(uint32_t seq_now, uint32_t seq_before, int64_t enqueued, int64_t run) {
  if (enqueued <= 0) return seq_now;
  const uint32_t per_iter = (seq_now - seq_before) / (uint32_t)enqueued;
  if (per_iter == 0 || run >= enqueued) return seq_now;  // nothing skipped: parity already alternates
  const uint32_t last_real = seq_before + per_iter * (uint32_t)std::max<int64_t>(run, 0);
  return (((seq_now + 1u) & 1u) == (last_real & 1u)) ? seq_now + 1u : seq_now;
}

}  // namespace kb
