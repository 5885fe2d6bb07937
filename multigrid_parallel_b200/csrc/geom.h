// geom.h -- device data layout of one grid level (shared by kernels and host).
//
// Colour-split storage.  A grid of ni x nj x nk points (k contiguous in the
// reference's natural layout, mg_3d.h:43-44) is stored as TWO arrays, one per
// colour c = (i+j+k)&1 (1 = red, 0 = black; mg_3d.h:669-693).  In row (i,j),
// with row parity s = (i+j)&1, colour c holds the points k = 2m + (c^s),
// m = 0,1,...  at
//
//     base + c*cs + (il*nj + j)*kh + m          (il = i - i0, local plane)
//
// kh (half-row pitch) is (nk+1)/2 rounded up to a multiple of 4 doubles
// (32-byte sectors, 16-byte vector alignment); pad entries are kept at 0.
// A plane of one colour is therefore one contiguous run of nj*kh doubles, and
// the six neighbours of a point of colour c all live in the OTHER colour's
// array: (i+-1,j,k) and (i,j+-1,k) at the same m, (i,j,k-1)/(i,j,k+1) at
// m+kp-1 / m+kp with kp = c^s.  A half-sweep thus reads one colour and writes
// the other: 12 B/DOF, all 128-bit and coalesced.
#pragma once

struct Geo {
    int ni, nj, nk;  // global extents of the level
    int li;          // planes stored locally (slab + halos); == ni on one GPU
    int i0;          // global i of local plane 0
    int kh;          // half-row pitch in doubles (multiple of 4)
    long long pj;    // plane pitch in doubles = nj*kh
    long long cs;    // colour stride in doubles (>= li*pj, multiple of 16)
};

static inline int mgb_half_pitch(int nk) { return (((nk + 1) / 2) + 3) & ~3; }

// guard (in doubles) kept zeroed in front of and behind every level array so
// that the +-1/+2 element reads of masked lanes stay inside the allocation
#define MGB_GUARD 64
