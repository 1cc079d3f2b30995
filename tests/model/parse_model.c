/* tests/model/parse_model.c -- sequential CPU model of the range-parallel parse used by the GPU match finders.
 * TEST / DESIGN INFRASTRUCTURE ONLY (never linked into the product): it lets parse variants (table widths,
 * sparse insertion, long-hash preference, lazy step, repeat probes, range width) be scored for output size with
 * the oracle's entropy stage before a kernel is written.  The model is exact for the GPU algorithm because the
 * per-range walks only read static tables: running the ranges one after the other gives the same records.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t rangeBytes;   /* bytes per walker range (68) */
    uint32_t shortBits;    /* short table bits */
    uint32_t shortMls;     /* short hash length in bytes: 4, 5, 6 */
    uint32_t longBits;     /* 0 = no long table */
    uint32_t longMls;      /* 7 or 8 */
    uint32_t insStride;    /* 1 = every position inserted, 2 = even positions only */
    uint32_t lazy;         /* 1 = a short match yields to a long match at p+1 */
    uint32_t repProbe;     /* N = after a match probe N positions (end+1 ..) for the same offset */
    uint32_t repCodes;     /* 1 = only repeat code 1 (round-1 rule), 3 = full 3-entry repeat history */
    uint32_t joinCont;     /* 1 = a kept record with litLen 0 and the previous offset extends the previous sequence */
    uint32_t minLong;      /* bytes a long candidate must verify (4 or 8) */
    uint32_t minShort;     /* bytes a short candidate must verify: 4, or shortMls (models tag equality) */
    uint32_t maxDist;      /* >0: candidates farther than this are dropped */
    uint32_t hist;         /* bytes of history before the block (positions [0,hist) are history) */
    uint32_t latest;       /* 1 = tables hold the latest position of the history part (hist only) */
    uint32_t tile;         /* dyn only: >0 = a position only sees insertions from earlier tiles of this many positions */
    uint32_t local;        /* tile only: 1 = plus the latest equal-hash position inside the same aligned 32-position group */
    uint32_t lag;          /* dyn only: candidates are at least this far back (exact latest otherwise) */
    uint32_t probeStride;  /* 2 = only even positions are probed (candidates exist at even positions only) */
    uint32_t both;         /* 1 = the walk extends the near and the far candidate and keeps the longer match */
    uint32_t dyn;          /* 1 = candidate = latest earlier position with the same hash (upper bound: a dynamic table) */
} pm_cfg;

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

static inline uint32_t hashN(uint64_t u, uint32_t mls, uint32_t bits) {
    /* zstd/hash.go style multiplicative hashes */
    static const uint64_t prime[9] = {0, 0, 0, 0, 2654435761ull, 889523592379ull, 227718039650203ull,
                                      58295818150454627ull, 0xcf1bbcdcb7a56463ull};
    if (mls == 4) return (uint32_t)((uint32_t)u * 2654435761u) >> (32 - bits);
    return (uint32_t)(((u << (64 - 8 * mls)) * prime[mls]) >> (64 - bits));
}

typedef struct { uint32_t s, len, dist; } rec_t;

/* src has n + hist bytes followed by >= 16 bytes of zero padding.  Output: triples (ll, ml-3, offsetValue), lits. */
int pm_parse(const uint8_t *src, uint32_t n, const pm_cfg *c, uint32_t *triples, uint8_t *lits, uint32_t *nlit_out) {
    const uint32_t H = c->hist, N = H + n;
    const uint32_t npos = N >= 8 ? N - 7 : 0;
    uint32_t *ES = malloc(sizeof(uint32_t) << c->shortBits);
    uint32_t *EL = c->longBits ? malloc(sizeof(uint32_t) << c->longBits) : NULL;
    memset(ES, 0xff, sizeof(uint32_t) << c->shortBits);
    if (EL) memset(EL, 0xff, sizeof(uint32_t) << c->longBits);
    /* earliest occurrence (block part); for the history part optionally the latest */
    for (uint32_t p = 0; p < npos; p += c->insStride) {
        uint64_t v = rd64(src + p);
        uint32_t hs = hashN(v, c->shortMls, c->shortBits);
        if (ES[hs] == 0xffffffffu || (c->latest && p < H)) ES[hs] = p;
        if (EL) {
            uint32_t hl = hashN(v, c->longMls, c->longBits);
            if (EL[hl] == 0xffffffffu || (c->latest && p < H)) EL[hl] = p;
        }
    }
    uint32_t *PS = NULL, *PL = NULL, *FS = NULL, *FLg = NULL;   /* F*: far candidates kept beside the near ones (both=1) */
    if (c->dyn) {
        PS = malloc(4 * (npos + 1)); PL = malloc(4 * (npos + 1)); FS = malloc(4 * (npos + 1)); FLg = malloc(4 * (npos + 1));
        memset(FS, 0xff, 4 * (npos + 1)); memset(FLg, 0xff, 4 * (npos + 1));
        memset(ES, 0xff, sizeof(uint32_t) << c->shortBits);
        if (EL) memset(EL, 0xff, sizeof(uint32_t) << c->longBits);
        const uint32_t TL = c->tile ? c->tile : 1;
        for (uint32_t t0 = 0; t0 < npos; t0 += TL) {
            const uint32_t t1 = (t0 + TL < npos) ? t0 + TL : npos;
            for (uint32_t p = t0; p < t1; p++) {
                uint64_t v = rd64(src + p);
                PS[p] = ES[hashN(v, c->shortMls, c->shortBits)];
                if (EL) PL[p] = EL[hashN(v, c->longMls, c->longBits)];
                FS[p] = PS[p]; if (EL) FLg[p] = PL[p];
                if (c->local) {   /* nearest earlier position of the same 32-group with the same hash */
                    for (uint32_t q = p; q-- > (p & ~31u);) {
                        uint64_t vq = rd64(src + q);
                        if (q >= t0 && hashN(vq, c->shortMls, c->shortBits) == hashN(v, c->shortMls, c->shortBits)) { PS[p] = q; break; }
                    }
                    if (EL) for (uint32_t q = p; q-- > (p & ~31u);) {
                        uint64_t vq = rd64(src + q);
                        if (q >= t0 && hashN(vq, c->longMls, c->longBits) == hashN(v, c->longMls, c->longBits)) { PL[p] = q; break; }
                    }
                }
            }
            if (c->local == 2) {
                /* insert: earliest position of this tile per hash (overwrites entries of older tiles) */
                for (uint32_t p = t1; p-- > t0;) {
                    if (p % c->insStride) continue;
                    uint64_t v = rd64(src + p);
                    ES[hashN(v, c->shortMls, c->shortBits)] = p;
                    if (EL) EL[hashN(v, c->longMls, c->longBits)] = p;
                }
                /* second probe: the tile's earliest occurrence, when it precedes p, replaces the far candidate */
                for (uint32_t p = t0; p < t1; p++) {
                    uint64_t v = rd64(src + p);
                    uint32_t q = ES[hashN(v, c->shortMls, c->shortBits)];
                    if (q < p && (c->minShort > 4 ? ((rd64(src + q) ^ v) << (64 - 8 * c->minShort)) == 0 : rd32(src + q) == (uint32_t)v)) PS[p] = q;
                    if (EL) { q = EL[hashN(v, c->longMls, c->longBits)]; if (q < p && (c->minLong == 8 ? rd64(src + q) == v : rd32(src + q) == (uint32_t)v)) PL[p] = q; }
                }
                continue;
            }
            for (uint32_t p = t0; p < t1; p++) {
                if (p % c->insStride) continue;
                if (c->lag) {   /* lagged insertion: position p - lag becomes visible when p is processed */
                    if (p < c->lag) continue;
                    uint64_t vv = rd64(src + p - c->lag);
                    ES[hashN(vv, c->shortMls, c->shortBits)] = p - c->lag;
                    if (EL) EL[hashN(vv, c->longMls, c->longBits)] = p - c->lag;
                    continue;
                }
                uint64_t v = rd64(src + p);
                ES[hashN(v, c->shortMls, c->shortBits)] = p;
                if (EL) EL[hashN(v, c->longMls, c->longBits)] = p;
            }
        }
    }
    const uint32_t RB = c->rangeBytes;
    const uint32_t nr = (n + RB - 1) / RB;
    rec_t *recs = malloc(sizeof(rec_t) * (n / 4 + 16));
    uint32_t *rfirst = malloc(sizeof(uint32_t) * (nr + 1));
    uint32_t nrec = 0;
    for (uint32_t r = 0; r < nr; r++) {
        rfirst[r] = nrec;
        const uint32_t b = H + r * RB, e = (b + RB < N) ? b + RB : N;
        const uint32_t pend = e < npos ? e : npos;
        uint32_t p = b, nextEmit = b, prevOff = 0, repLeft = 0;
        while (p < pend) {
            uint64_t v = rd64(src + p);
            uint32_t cand = 0xffffffffu;
            if (c->probeStride == 2 && (p & 1)) { p++; continue; }
            if (repLeft && prevOff && p >= prevOff && rd32(src + p - prevOff) == (uint32_t)v) cand = p - prevOff;
            if (repLeft) repLeft--;
            if (cand == 0xffffffffu && EL) {
                uint32_t q = c->dyn ? PL[p] : EL[hashN(v, c->longMls, c->longBits)];
                if (q < p && (c->minLong == 8 ? rd64(src + q) == v : rd32(src + q) == (uint32_t)v)) cand = q;
            }
            if (cand == 0xffffffffu) {
                uint32_t q = c->dyn ? PS[p] : ES[hashN(v, c->shortMls, c->shortBits)];
                if (q < p && (c->minShort > 4 ? ((rd64(src + q) ^ v) << (64 - 8 * c->minShort)) == 0 : rd32(src + q) == (uint32_t)v)) {
                    cand = q;
                    if (c->lazy && EL && p + 1 < pend) {
                        uint64_t v1 = rd64(src + p + 1);
                        uint32_t q1 = c->dyn ? PL[p + 1] : EL[hashN(v1, c->longMls, c->longBits)];
                        if (q1 < p + 1 && (c->minLong == 8 ? rd64(src + q1) == v1 : rd32(src + q1) == (uint32_t)v1)) { p = p + 1; cand = q1; }
                    }
                }
            }
            if (cand != 0xffffffffu && c->maxDist && p - cand > c->maxDist) cand = 0xffffffffu;
            if (cand == 0xffffffffu) { p++; continue; }
            uint32_t len = 4;
            while (p + len < N && src[p + len] == src[cand + len]) len++;
            if (c->both && c->dyn) {
                uint32_t alts[2] = {FS[p], EL ? FLg[p] : 0xffffffffu};
                for (int a = 0; a < 2; a++) {
                    uint32_t q = alts[a];
                    if (q == 0xffffffffu || q >= p || q == cand || rd32(src + q) != (uint32_t)v) continue;
                    if (c->maxDist && p - q > c->maxDist) continue;
                    uint32_t l2 = 4;
                    while (p + l2 < N && src[p + l2] == src[q + l2]) l2++;
                    if (l2 > len + 1) { len = l2; cand = q; }
                }
            }
            uint32_t s = p, t = cand;
            while (s > nextEmit && t > 0 && src[s - 1] == src[t - 1]) { s--; t--; len++; }
            recs[nrec].s = s; recs[nrec].len = len; recs[nrec].dist = p - cand; nrec++;
            p = s + len; nextEmit = p; prevOff = recs[nrec - 1].dist;
            repLeft = c->repProbe ? c->repProbe + 1 : 0;   /* the probe at the end position itself always fails (the match would have been longer) */
        }
    }
    rfirst[nr] = nrec;
    /* merge: prefix max of ends, trim */
    uint32_t R = H, nk = 0;
    rec_t *kept = malloc(sizeof(rec_t) * (nrec + 1));
    for (uint32_t r = 0; r < nr; r++) {
        uint32_t lastE = 0;
        for (uint32_t j = rfirst[r]; j < rfirst[r + 1]; j++) {
            rec_t x = recs[j];
            uint32_t e0 = x.s + x.len;
            lastE = e0;   /* the range's last record end (its own records are ordered) */
            if (e0 <= R) continue;
            uint32_t s2 = x.s > R ? x.s : R, l2 = e0 - s2;
            if (l2 < 4) continue;
            kept[nk].s = s2; kept[nk].len = l2; kept[nk].dist = x.dist; nk++;
        }
        if (rfirst[r + 1] > rfirst[r] && lastE > R) R = lastE;
    }
    /* continuation join */
    if (c->joinCont) {
        uint32_t m = 0;
        for (uint32_t i = 0; i < nk; i++) {
            if (m && kept[i].s == kept[m - 1].s + kept[m - 1].len && kept[i].dist == kept[m - 1].dist) kept[m - 1].len += kept[i].len;
            else kept[m++] = kept[i];
        }
        nk = m;
    }
    /* sequences + repeat codes */
    uint32_t rep[3] = {1, 4, 8};
    uint32_t prevE = H, nl = 0;
    for (uint32_t i = 0; i < nk; i++) {
        uint32_t ll = kept[i].s - prevE, d = kept[i].dist, ofv = d + 3;
        memcpy(lits + nl, src + prevE, ll); nl += ll;
        if (c->repCodes == 1) {
            if (i > 0 && ll > 0 && d == rep[0]) ofv = 1;
            rep[0] = d;
        } else {
            if (ll > 0) {
                if (d == rep[0]) ofv = 1;
                else if (d == rep[1]) { ofv = 2; rep[1] = rep[0]; rep[0] = d; }
                else if (d == rep[2]) { ofv = 3; rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = d; }
                else { rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = d; }
            } else {
                if (d == rep[1]) { ofv = 1; rep[1] = rep[0]; rep[0] = d; }
                else if (d == rep[2]) { ofv = 2; rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = d; }
                else if (d == rep[0] - 1 && rep[0] > 1) { ofv = 3; rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = d; }
                else { rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = d; }
            }
        }
        triples[3 * i] = ll; triples[3 * i + 1] = kept[i].len - 3; triples[3 * i + 2] = ofv;
        prevE = kept[i].s + kept[i].len;
    }
    memcpy(lits + nl, src + prevE, N - prevE); nl += N - prevE;
    *nlit_out = nl;
    free(PS); free(PL); free(FS); free(FLg); free(ES); free(EL); free(recs); free(rfirst); free(kept);
    return (int)nk;
}
