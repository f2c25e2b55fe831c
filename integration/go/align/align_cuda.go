//go:build cuda

// Package align — CUDA backend for the pairwise DP entry points.
//
// Drop this file (and align_cuda_engine.go) into gonomics' align/ directory next to the existing
// sources, rename the pure-Go bodies' file guards to `//go:build !cuda`, and build with
//
//	CGO_CFLAGS="-I/path/to/repo/include" CGO_LDFLAGS="-L/path/to/repo/gonomics_b200 -lgnxalign" go build -tags cuda ./...
//
// Signatures, results and panics are those of the reference functions; only the work moves to the
// GPU through the C ABI in include/gnxalign.h.  NOT COMPILED in the build image (no Go toolchain there).
package align

/*
#cgo LDFLAGS: -lgnxalign
#include <stdlib.h>
#include "gnxalign.h"
*/
import "C"

import (
	"log"
	"runtime"
	"sync"
	"unsafe"

	"github.com/vertgenlab/gonomics/dna"
)

// one gnx_ctx per OS thread that calls in (contexts are not thread-safe); pooled like the
// per-worker scratch matrices of cmd/gsw.
var ctxPool = sync.Pool{New: func() any {
	c := C.gnx_create(0, 0)
	if c == nil {
		log.Panicf("gnxalign: %s", C.GoString(C.gnx_last_error(nil)))
	}
	return c
}}

func flatten(scores [][]int64) ([]int64, int) {
	dim := len(scores)
	flat := make([]int64, dim*dim) // cgo forbids passing Go pointers to Go pointers
	for i := range scores {
		copy(flat[i*dim:], scores[i])
	}
	return flat, dim
}

func basePtr(s []dna.Base) *C.uint8_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&s[0])) // dna.Base is a byte: no copy
}

func check(ctx *C.gnx_ctx, rc C.int) {
	switch rc {
	case C.GNX_OK:
	case C.GNX_EBASE: // the reference indexes scores[alpha[i]][beta[j]] out of range
		panic("runtime error: index out of range (base >= len(scores))")
	case C.GNX_ECHUNK:
		log.Fatalf("Error: sequence length should be a multiple of chunkSize\n")
	default:
		log.Panicf("gnxalign: %s", C.GoString(C.gnx_last_error(ctx)))
	}
}

// alignOne runs a single pair; mode 0 = AffineGap_highMem, 1 = AffineGapLocal, 2 = ConstGap_highMem.
func alignOne(alpha, beta []dna.Base, scores [][]int64, gapOpen, gapExtend int64, mode int) (int64, []Cigar) {
	ctx := ctxPool.Get().(*C.gnx_ctx)
	defer ctxPool.Put(ctx)
	flat, dim := flatten(scores)
	aoff := [2]C.int64_t{0, C.int64_t(len(alpha))}
	boff := [2]C.int64_t{0, C.int64_t(len(beta))}
	var score C.int64_t
	var coff [2]C.int64_t
	route := make([]Cigar, len(alpha)+len(beta)+1) // align.Cigar == gnx_cigar (16 B), written in place
	var rc C.int
	if mode == 2 {
		rc = C.gnx_const_batch(ctx, basePtr(alpha), &aoff[0], basePtr(beta), &boff[0], 1,
			(*C.int64_t)(unsafe.Pointer(&flat[0])), C.int(dim), C.int64_t(gapOpen), 1,
			&score, (*C.gnx_cigar)(unsafe.Pointer(&route[0])), &coff[0], C.int64_t(len(route)))
	} else {
		rc = C.gnx_affine_batch(ctx, basePtr(alpha), &aoff[0], basePtr(beta), &boff[0], 1,
			(*C.int64_t)(unsafe.Pointer(&flat[0])), C.int(dim), C.int64_t(gapOpen), C.int64_t(gapExtend),
			C.int(mode), 1, &score, (*C.gnx_cigar)(unsafe.Pointer(&route[0])), &coff[0], C.int64_t(len(route)))
	}
	runtime.KeepAlive(flat)
	check(ctx, rc)
	return int64(score), route[:coff[1]]
}

// AffineGap_highMem: see align/affineGap_highMem.go:99.
func AffineGap_highMem(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapOpen int64, gapExtend int64) (int64, []Cigar) {
	return alignOne(alpha, beta, scores, gapOpen, gapExtend, 0)
}

// AffineGapLocal: see align/affineGap_highMem.go:105.
func AffineGapLocal(target []dna.Base, query []dna.Base, scores [][]int64, gapOpen int64, gapExtend int64) (int64, []Cigar) {
	return alignOne(target, query, scores, gapOpen, gapExtend, 1)
}

// AffineGap: see align/affineGap.go:59 (checker size 10000 x 10000: one board for len <= 10000).
func AffineGap(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapOpen int64, gapExtend int64) (int64, []Cigar) {
	return AffineGap_customizeCheckersize(alpha, beta, scores, gapOpen, gapExtend, 10000, 10000)
}

// AffineGap_customizeCheckersize: see align/affineGap.go:73.  The checkerboard only bounds the
// reference's memory; the GPU keeps the packed trace in HBM, so the sizes are accepted and ignored.
func AffineGap_customizeCheckersize(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapOpen int64, gapExtend int64, checkersize_i int, checkersize_j int) (int64, []Cigar) {
	if len(alpha) == 0 || len(beta) == 0 {
		panic("runtime error: index out of range") // what the reference does on an empty input
	}
	return alignOne(alpha, beta, scores, gapOpen, gapExtend, 0)
}

// ConstGap_highMem: see align/constGap_highMem.go:11.
func ConstGap_highMem(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapPen int64) (int64, []Cigar) {
	return alignOne(alpha, beta, scores, gapPen, 0, 2)
}

// ConstGap: see align/constGap.go:13.
func ConstGap(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapPen int64) (int64, []Cigar) {
	return ConstGap_customizeCheckersize(alpha, beta, scores, gapPen, 10000, 10000)
}

// ConstGap_customizeCheckersize: see align/constGap.go:73.
func ConstGap_customizeCheckersize(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapPen int64, checkersize_i int, checkersize_j int) (int64, []Cigar) {
	if len(alpha) == 0 || len(beta) == 0 {
		panic("runtime error: index out of range")
	}
	return alignOne(alpha, beta, scores, gapPen, 0, 2)
}

// AffineGapBatch is the performant boundary (new, additive): one GPU call for a slice of pairs.
// mode: 0 global (AffineGap_highMem), 1 free end gaps (AffineGapLocal).
func AffineGapBatch(targets, queries [][]dna.Base, scores [][]int64, gapOpen, gapExtend int64, mode int) ([]int64, [][]Cigar) {
	n := len(targets)
	ctx := ctxPool.Get().(*C.gnx_ctx)
	defer ctxPool.Put(ctx)
	flat, dim := flatten(scores)
	aoff := make([]int64, n+1)
	boff := make([]int64, n+1)
	for i := 0; i < n; i++ {
		aoff[i+1] = aoff[i] + int64(len(targets[i]))
		boff[i+1] = boff[i] + int64(len(queries[i]))
	}
	acat := make([]dna.Base, aoff[n]) // one concatenation copy; callers that already hold
	bcat := make([]dna.Base, boff[n]) // concatenated reads can call the C ABI directly
	for i := 0; i < n; i++ {
		copy(acat[aoff[i]:], targets[i])
		copy(bcat[boff[i]:], queries[i])
	}
	out := make([]int64, n)
	coff := make([]int64, n+1)
	cig := make([]Cigar, 16*n+64)
	rc := C.gnx_affine_batch(ctx, basePtr(acat), (*C.int64_t)(unsafe.Pointer(&aoff[0])), basePtr(bcat),
		(*C.int64_t)(unsafe.Pointer(&boff[0])), C.int64_t(n), (*C.int64_t)(unsafe.Pointer(&flat[0])), C.int(dim),
		C.int64_t(gapOpen), C.int64_t(gapExtend), C.int(mode), 1, (*C.int64_t)(unsafe.Pointer(&out[0])),
		(*C.gnx_cigar)(unsafe.Pointer(&cig[0])), (*C.int64_t)(unsafe.Pointer(&coff[0])), C.int64_t(len(cig)))
	if rc == C.GNX_ECAP { // offsets are valid: grow and fetch the retained cigars
		cig = make([]Cigar, coff[n])
		rc = C.gnx_copy_last_cigars(ctx, (*C.gnx_cigar)(unsafe.Pointer(&cig[0])), C.int64_t(len(cig)))
	}
	check(ctx, rc)
	routes := make([][]Cigar, n)
	for i := 0; i < n; i++ {
		routes[i] = cig[coff[i]:coff[i+1]:coff[i+1]]
	}
	return out, routes
}
