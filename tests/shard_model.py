"""CPU model (Python ints, test infrastructure) of the ONE-PROOF-OVER-G-RANKS polynomial stage that
zksnark-rs_b200/csrc/shard.cu runs on the device: the NTT's outer dimension sharded over G = 2^lg ranks,
one all-to-all per distributed transform, log G butterfly stages instead of an O(G) sum.

Index conventions (n = G*m, m = G*q):
  layout D (decimated):  rank r owns x[r + G*j2], j2 < m                      (local index j2)
  layout S (strided blocks): rank r owns x[(r*q + t) + m*k1], t < q, k1 < G   (local index s = k1*q + t)
  DIT-distributed transform  D -> S:  local size-m transform, all-to-all, twiddle w^(g*k2) and a size-G transform
  DIF-distributed transform  S -> D:  size-G transform and twiddle w^(j2*k1), all-to-all, local size-m transform
The proof's pipeline  evaluations (D) -iNTT-> coefficients (S) -coset NTT-> evaluations (D) -iNTT-> coefficients (S)
therefore needs three exchanges (3, 2 and 1 vectors) and leaves u_sum, v_sum, h in layout S, which is also how the CRS
vectors xi / xi_t are sharded (any partition of an MSM's terms gives the same sum).

`exchange(send)`: send[d] = what this rank sends to rank d; returns recv with recv[g] = what rank g sent to this rank.
"""

from oracle import poly
from oracle.fields import FR

P = FR.p


def omega(log_n):
    return pow(5, (P - 1) >> log_n, P)


def s_to_j(s, r, G, n):
    m = n // G
    q = m // G
    k1, t = divmod(s, q)
    return (r * q + t) + m * k1


def small_dft(x, w):
    """size-len(x) transform X[k] = sum_j x[j] w^(j k) by radix-2 DIF stages (what the device does in registers)."""
    G = len(x)
    x = list(x)
    half = G // 2
    stride = 1
    while half >= 1:
        for b in range(0, G, 2 * half):
            for i in range(half):
                a, c = x[b + i], x[b + i + half]
                x[b + i] = (a + c) % P
                x[b + i + half] = (a - c) * pow(w, i * stride, P) % P
        half //= 2
        stride *= 2
    lg = G.bit_length() - 1
    out = [0] * G
    for p in range(G):
        out[int(format(p, f"0{lg}b")[::-1], 2) if lg else 0] = x[p]
    return out


def dit_distributed(local, r, G, n, w, exchange):
    """layout D -> layout S; X[k] = sum_j x[j] w^(j k) (no scaling)."""
    m = n // G
    q = m // G
    Y = poly.ntt_fast(FR, local, pow(w, G, P))                      # Y_r[k2], k2 < m
    recv = exchange([Y[d * q:(d + 1) * q] for d in range(G)])        # recv[g][t] = Y_g[r*q + t]
    out = [0] * m
    wG = pow(w, m, P)
    for t in range(q):
        k2 = r * q + t
        y = [recv[g][t] * pow(w, g * k2, P) % P for g in range(G)]
        X = small_dft(y, wG)                                         # X[k1] = sum_g wG^(g k1) y[g]
        for k1 in range(G):
            out[k1 * q + t] = X[k1]
    return out


def dif_distributed(local_S, r, G, n, w, exchange):
    """layout S -> layout D; X[k] = sum_j x[j] w^(j k)."""
    m = n // G
    q = m // G
    wG = pow(w, m, P)
    send = [[0] * q for _ in range(G)]
    for t in range(q):
        j2 = r * q + t
        Z = small_dft([local_S[j1 * q + t] for j1 in range(G)], wG)  # Z[k1] = sum_j1 wG^(j1 k1) x[j2 + m j1]
        for k1 in range(G):
            send[k1][t] = Z[k1] * pow(w, j2 * k1, P) % P
    recv = exchange(send)                                            # recv[g][t] = W_r[g*q + t]
    W = [v for g in range(G) for v in recv[g]]
    return poly.ntt_fast(FR, W, pow(w, G, P))                        # X[r + G*k2], k2 < m


def poly_stage_rank(A_loc, B_loc, r, G, n, exchange):
    """A_loc[j2] = A_{r + G j2} etc. (layout D).  Returns (u, v, h) in layout S (h[n-1] = 0 included)."""
    log_n = n.bit_length() - 1
    w = omega(log_n)
    winv = pow(w, -1, P)
    g2n = omega(log_n + 1)
    ninv = pow(n, -1, P)
    AB = [a * b % P for a, b in zip(A_loc, B_loc)]
    u = [x * ninv % P for x in dit_distributed(A_loc, r, G, n, winv, exchange)]
    v = [x * ninv % P for x in dit_distributed(B_loc, r, G, n, winv, exchange)]
    c = [x * ninv % P for x in dit_distributed(AB, r, G, n, winv, exchange)]
    m = n // G
    cos = [pow(g2n, s_to_j(s, r, G, n), P) for s in range(m)]
    U = dif_distributed([a * b % P for a, b in zip(u, cos)], r, G, n, w, exchange)
    V = dif_distributed([a * b % P for a, b in zip(v, cos)], r, G, n, w, exchange)
    d = [x * ninv % P for x in dit_distributed([a * b % P for a, b in zip(U, V)], r, G, n, winv, exchange)]
    inv2 = pow(2, -1, P)
    h = [(cc - dd * pow(ci, -1, P)) * inv2 % P for cc, dd, ci in zip(c, d, cos)]
    return u, v, h


def run_all_ranks(fn, G):
    """fn(r, exchange) for every rank r < G on its own thread; `exchange` is an all-to-all through a shared mailbox."""
    import threading
    box = [[None] * G for _ in range(G)]
    bar = threading.Barrier(G)
    results, errors = [None] * G, []

    def make_exchange(r):
        def ex(send):
            for d in range(G):
                box[d][r] = send[d]
            bar.wait()
            recv = list(box[r])
            bar.wait()
            return recv
        return ex

    def work(r):
        try:
            results[r] = fn(r, make_exchange(r))
        except Exception as e:  # pragma: no cover
            errors.append(e)
            bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(G)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    return results
