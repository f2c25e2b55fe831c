//go:build cuda

// Package align — CUDA backend for the pairwise DP entry points.
//
// Drop this file, align_cuda_engine.go and align_cuda_multi.go into gonomics' align/ directory next to the existing
// sources, put `//go:build !cuda` on the pure-Go bodies of the same names (see align_nocuda.go.txt for the list), and build with
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
	"unsafe"

	"github.com/vertgenlab/gonomics/dna"
)

// Contexts are not thread-safe and own device memory, streams and page-locked buffers, so they live in a BOUNDED
// free list (a sync.Pool would drop idle contexts at GC without ever calling gnx_destroy).  A goroutine takes one
// for the duration of a call; when all are out, the caller blocks -- the GPU is the bottleneck then anyway.
// Every context gets an explicit workspace so that concurrent contexts do not each claim 2/3 of the device.
const (
	maxContexts    = 8
	workspaceBytes = 8 << 30 // traceback matrices of the chunks in flight; long pairs use per-warp scratch instead
)

var (
	ctxFree    = make(chan *C.gnx_ctx, maxContexts)
	ctxCreated = make(chan struct{}, maxContexts) // one token per context ever created
)

// Device selects the CUDA device new contexts are created on (set before the first call; default 0).  For several
// GPUs behind one call use MultiAffineGapBatch (align_cuda_multi.go).
var Device = 0

func getCtx() *C.gnx_ctx {
	select {
	case c := <-ctxFree:
		return c
	default:
	}
	select {
	case ctxCreated <- struct{}{}: // below the bound: create one more
		c := C.gnx_create(C.int(Device), C.size_t(workspaceBytes))
		if c == nil {
			<-ctxCreated
			log.Panicf("gnxalign: %s", C.GoString(C.gnx_last_error(nil)))
		}
		return c
	case c := <-ctxFree: // at the bound: wait for one to come back
		return c
	}
}

func putCtx(c *C.gnx_ctx) { ctxFree <- c }

// Shutdown destroys every idle context (call at process exit, after the last alignment).
func Shutdown() {
	for {
		select {
		case c := <-ctxFree:
			C.gnx_destroy(c)
			<-ctxCreated
		default:
			return
		}
	}
}

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

func i64Ptr(s []int64) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}

func cigPtr(s []Cigar) *C.gnx_cigar {
	if len(s) == 0 {
		return nil
	}
	return (*C.gnx_cigar)(unsafe.Pointer(&s[0])) // align.Cigar == gnx_cigar (16 B)
}

func check(ctx *C.gnx_ctx, rc C.int) {
	switch rc {
	case C.GNX_OK:
	case C.GNX_EBASE: // the reference indexes scores[alpha[i]][beta[j]] out of range
		panic("runtime error: index out of range (base >= len(scores))")
	case C.GNX_ECHUNK:
		log.Fatalf("Error: sequence length should be a multiple of chunkSize\n")
	case C.GNX_EDIVZERO:
		panic("runtime error: integer divide by zero") // scoreColumnMatch on an all-gap column pair
	default:
		log.Panicf("gnxalign: %s", C.GoString(C.gnx_last_error(ctx)))
	}
}

// alignOne runs a single pair; mode 0 = AffineGap_highMem, 1 = AffineGapLocal, 2 = ConstGap_highMem.
func alignOne(alpha, beta []dna.Base, scores [][]int64, gapOpen, gapExtend int64, mode int) (int64, []Cigar) {
	ctx := getCtx()
	defer putCtx(ctx)
	flat, dim := flatten(scores)
	aoff := [2]C.int64_t{0, C.int64_t(len(alpha))}
	boff := [2]C.int64_t{0, C.int64_t(len(beta))}
	var score C.int64_t
	var coff [2]C.int64_t
	route := make([]Cigar, len(alpha)+len(beta)+1) // written in place by the library
	var rc C.int
	if mode == 2 {
		rc = C.gnx_const_batch(ctx, basePtr(alpha), &aoff[0], basePtr(beta), &boff[0], 1,
			i64Ptr(flat), C.int(dim), C.int64_t(gapOpen), 1,
			&score, cigPtr(route), &coff[0], C.int64_t(len(route)))
	} else {
		rc = C.gnx_affine_batch(ctx, basePtr(alpha), &aoff[0], basePtr(beta), &boff[0], 1,
			i64Ptr(flat), C.int(dim), C.int64_t(gapOpen), C.int64_t(gapExtend),
			C.int(mode), 1, &score, cigPtr(route), &coff[0], C.int64_t(len(route)))
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

// requireOneBoard: for inputs that fit one checkerboard the low-memory drivers equal the high-memory result.  Past
// one board the reference's stitching has defects (SURVEY.md 8a) that the GPU path does not reproduce; instead of
// returning a silently different cigar the call fails -- callers that want the high-memory alignment of a longer
// input call AffineGap_highMem / ConstGap_highMem.
func requireOneBoard(alpha, beta []dna.Base, ci, cj int, what string) {
	if len(alpha) == 0 || len(beta) == 0 {
		panic("runtime error: index out of range") // what the reference does on an empty input
	}
	if len(alpha) > ci || len(beta) > cj {
		log.Panicf("%s: %d x %d spans more than one %d x %d checkerboard; the multi-board stitching of the pure-Go "+
			"path is not reproduced on the GPU -- call the _highMem function", what, len(alpha), len(beta), ci, cj)
	}
}

// AffineGap_customizeCheckersize: see align/affineGap.go:73.
func AffineGap_customizeCheckersize(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapOpen int64, gapExtend int64, checkersize_i int, checkersize_j int) (int64, []Cigar) {
	requireOneBoard(alpha, beta, checkersize_i, checkersize_j, "AffineGap_customizeCheckersize")
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
	requireOneBoard(alpha, beta, checkersize_i, checkersize_j, "ConstGap_customizeCheckersize")
	return alignOne(alpha, beta, scores, gapPen, 0, 2)
}

// AffineGapChunk: see align/affineGap_highMem.go:227.
func AffineGapChunk(alpha []dna.Base, beta []dna.Base, scores [][]int64, gapOpen int64, gapExtend int64, chunkSize int64) (int64, []Cigar) {
	ctx := getCtx()
	defer putCtx(ctx)
	flat, dim := flatten(scores)
	aoff := [2]C.int64_t{0, C.int64_t(len(alpha))}
	boff := [2]C.int64_t{0, C.int64_t(len(beta))}
	var score C.int64_t
	var coff [2]C.int64_t
	route := make([]Cigar, len(alpha)+len(beta)+1)
	rc := C.gnx_affine_chunk_batch(ctx, basePtr(alpha), &aoff[0], basePtr(beta), &boff[0], 1, i64Ptr(flat), C.int(dim),
		C.int64_t(gapOpen), C.int64_t(gapExtend), C.int64_t(chunkSize), &score, cigPtr(route), &coff[0], C.int64_t(len(route)))
	runtime.KeepAlive(flat)
	check(ctx, rc)
	return int64(score), route[:coff[1]]
}

// concat flattens a slice of sequences into the library's (bytes, offsets) batch form.
func concat(seqs [][]dna.Base) ([]dna.Base, []int64) {
	off := make([]int64, len(seqs)+1)
	for i, s := range seqs {
		off[i+1] = off[i] + int64(len(s))
	}
	cat := make([]dna.Base, off[len(seqs)])
	for i, s := range seqs {
		copy(cat[off[i]:], s)
	}
	return cat, off
}

// AffineGapBatch is the performant boundary (new, additive): one GPU call for a slice of pairs.
// mode: 0 global (AffineGap_highMem), 1 free end gaps (AffineGapLocal).  Callers that already hold their reads
// concatenated (or in dnaTwoBit form) should call AffineGapBatchCat / AffineGapBatchTwoBit and skip the copy.
func AffineGapBatch(targets, queries [][]dna.Base, scores [][]int64, gapOpen, gapExtend int64, mode int) ([]int64, [][]Cigar) {
	n := len(targets)
	if n == 0 {
		return nil, nil
	}
	acat, aoff := concat(targets)
	bcat, boff := concat(queries)
	out, coff, cig := AffineGapBatchCat(acat, aoff, bcat, boff, scores, gapOpen, gapExtend, mode)
	routes := make([][]Cigar, n)
	for i := 0; i < n; i++ {
		routes[i] = cig[coff[i]:coff[i+1]:coff[i+1]]
	}
	return out, routes
}

// AffineGapBatchCat: pair p is acat[aoff[p]:aoff[p+1]] against bcat[boff[p]:boff[p+1]]; returns scores, cigar
// offsets (n+1) and the cigars of all pairs back to back.
func AffineGapBatchCat(acat []dna.Base, aoff []int64, bcat []dna.Base, boff []int64, scores [][]int64, gapOpen, gapExtend int64, mode int) ([]int64, []int64, []Cigar) {
	n := len(aoff) - 1
	if n <= 0 {
		return nil, []int64{0}, nil
	}
	ctx := getCtx()
	defer putCtx(ctx)
	flat, dim := flatten(scores)
	out := make([]int64, n)
	coff := make([]int64, n+1)
	cig := make([]Cigar, 16*n+64)
	rc := C.gnx_affine_batch(ctx, basePtr(acat), i64Ptr(aoff), basePtr(bcat), i64Ptr(boff), C.int64_t(n), i64Ptr(flat), C.int(dim),
		C.int64_t(gapOpen), C.int64_t(gapExtend), C.int(mode), 1, i64Ptr(out), cigPtr(cig), i64Ptr(coff), C.int64_t(len(cig)))
	if rc == C.GNX_ECAP { // offsets are valid: grow and fetch the retained cigars
		cig = make([]Cigar, coff[n])
		rc = C.gnx_copy_last_cigars(ctx, cigPtr(cig), C.int64_t(len(cig)))
	}
	runtime.KeepAlive(flat)
	check(ctx, rc)
	return out, coff, cig[:coff[n]]
}

// AffineGapBatchTwoBit aligns a batch held in dnaTwoBit form (gnx_affine_batch_twobit): words is the concatenation
// of the TwoBit.Seq slices (sequence p occupies ceil(Len/32) words), lens the TwoBit.Len values; pass nil lens and
// the common length for a uniform batch.  A quarter of the bytes cross PCIe.
func AffineGapBatchTwoBit(aWords []uint64, aLens []int64, aUniform int64, bWords []uint64, bLens []int64, bUniform int64, nPairs int,
	scores [][]int64, gapOpen, gapExtend int64, mode int) ([]int64, []int64, []Cigar) {
	if nPairs <= 0 {
		return nil, []int64{0}, nil
	}
	ctx := getCtx()
	defer putCtx(ctx)
	flat, dim := flatten(scores)
	out := make([]int64, nPairs)
	coff := make([]int64, nPairs+1)
	cig := make([]Cigar, 16*nPairs+64)
	rc := C.gnx_affine_batch_twobit(ctx, (*C.uint64_t)(unsafe.Pointer(&aWords[0])), i64Ptr(aLens), C.int64_t(aUniform),
		(*C.uint64_t)(unsafe.Pointer(&bWords[0])), i64Ptr(bLens), C.int64_t(bUniform), C.int64_t(nPairs), i64Ptr(flat), C.int(dim),
		C.int64_t(gapOpen), C.int64_t(gapExtend), C.int(mode), 1, i64Ptr(out), cigPtr(cig), i64Ptr(coff), C.int64_t(len(cig)))
	if rc == C.GNX_ECAP {
		cig = make([]Cigar, coff[nPairs])
		rc = C.gnx_copy_last_cigars(ctx, cigPtr(cig), C.int64_t(len(cig)))
	}
	runtime.KeepAlive(flat)
	check(ctx, rc)
	return out, coff, cig[:coff[nPairs]]
}
