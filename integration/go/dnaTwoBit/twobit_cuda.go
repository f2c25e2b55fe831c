//go:build cuda

// Package dnaTwoBit — CUDA backend for the 2-bit encoding and the perfect-match counters
// (dna/dnaTwoBit/dnaTwoBit.go:59-78, rainbow.go:8-25, perfectAlign.go:10-85).
//
// Additive: the reference's TwoBit{Seq []uint64; Len int} stays the host type; a TwoBitSet is a batch of
// TwoBit sequences resident on the GPU (a genome's nodes, a batch of reads), uploaded and packed once
// (gnx_twobit_new) and then queried in batches.  NOT COMPILED in the build image (no Go toolchain there).
//
//	CGO_CFLAGS="-I/path/to/repo/include" CGO_LDFLAGS="-L/path/to/repo/gonomics_b200 -lgnxalign" go build -tags cuda ./...
package dnaTwoBit

/*
#cgo LDFLAGS: -lgnxalign
#include "gnxalign.h"
*/
import "C"

import (
	"log"
	"runtime"
	"unsafe"

	"github.com/vertgenlab/gonomics/dna"
)

// TwoBitSet owns a gnx_twobit handle; Free it (or let the finalizer do so) when done.
type TwoBitSet struct {
	ctx *C.gnx_ctx
	h   *C.gnx_twobit
	N   int
}

func check(ctx *C.gnx_ctx, rc C.int) {
	switch rc {
	case C.GNX_OK:
	case C.GNX_EOFFSET: // perfectAlign.go:24-26, :63-65
		log.Fatalf("Error: Different offsets when comparing sequences\n")
	case C.GNX_EINDEX:
		panic("runtime error: index out of range")
	default:
		log.Panicf("gnxalign: %s", C.GoString(C.gnx_last_error(ctx)))
	}
}

// NewTwoBitSet packs every sequence like NewTwoBit (lead = 0) or like NewTwoBitRainbow(seq)[lead].
func NewTwoBitSet(ctx unsafe.Pointer, seqs [][]dna.Base, lead int) *TwoBitSet {
	c := (*C.gnx_ctx)(ctx)
	off := make([]int64, len(seqs)+1)
	for i, s := range seqs {
		off[i+1] = off[i] + int64(len(s))
	}
	cat := make([]dna.Base, off[len(seqs)]+1)
	for i, s := range seqs {
		copy(cat[off[i]:], s)
	}
	set := &TwoBitSet{ctx: c, N: len(seqs)}
	check(c, C.gnx_twobit_new(c, (*C.uint8_t)(unsafe.Pointer(&cat[0])), (*C.int64_t)(unsafe.Pointer(&off[0])),
		C.int64_t(len(seqs)), C.int(lead), &set.h))
	runtime.SetFinalizer(set, func(s *TwoBitSet) { s.Free() })
	return set
}

func (s *TwoBitSet) Free() {
	if s.h != nil {
		C.gnx_twobit_free(s.h)
		s.h = nil
	}
}

// Download returns the reference's TwoBit structs (Seq words and Len) of every sequence in the set.
func (s *TwoBitSet) Download() []TwoBit {
	var n, words C.int64_t
	C.gnx_twobit_info(s.h, &n, &words)
	seq := make([]uint64, int(words)+1)
	woff := make([]int64, int(n)+1)
	lens := make([]int64, int(n)+1)
	check(s.ctx, C.gnx_twobit_download(s.ctx, s.h, (*C.uint64_t)(unsafe.Pointer(&seq[0])),
		(*C.int64_t)(unsafe.Pointer(&woff[0])), (*C.int64_t)(unsafe.Pointer(&lens[0]))))
	out := make([]TwoBit, int(n))
	for i := range out {
		out[i] = TwoBit{Seq: seq[woff[i]:woff[i+1]:woff[i+1]], Len: int(lens[i])}
	}
	return out
}

// CountMatchesBatch runs CountRightMatches (left = false) or CountLeftMatches (left = true) for a list of
// (one[qOne[k]], startOne[k], two[qTwo[k]], startTwo[k]) queries in one GPU call.
func CountMatchesBatch(left bool, one, two *TwoBitSet, qOne, startOne, qTwo, startTwo []int64) []int64 {
	n := len(qOne)
	out := make([]int64, n+1)
	dir := C.int(C.GNX_MATCH_RIGHT)
	if left {
		dir = C.GNX_MATCH_LEFT
	}
	if n == 0 {
		return out[:0]
	}
	p := func(s []int64) *C.int64_t { return (*C.int64_t)(unsafe.Pointer(&s[0])) }
	check(one.ctx, C.gnx_twobit_count_matches(one.ctx, dir, one.h, two.h, p(qOne), p(startOne), p(qTwo), p(startTwo),
		C.int64_t(n), p(out)))
	return out[:n]
}
