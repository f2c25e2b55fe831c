//go:build cuda

// Package genomeGraph — CUDA backend for the two hot steps of cmd/gsw's seed-and-extend on a linear
// reference (nodes without edges): the perfect-match seed enumeration (seedMapMemPool, search.go:567-602)
// and the linear-gap extension DPs (LeftDynamicAln / RightDynamicAln, search.go:234-321).
//
// Both are batch-shaped: a worker collects the reads of a block, gets every read's seeds in one call,
// forms the (target window, read flank) pairs of the seeds it decides to extend, and gets every route in
// one call.  The per-read control flow (seedCouldBeBetter early exit, soft clips, flags: toGiraf.go:17-72)
// stays in Go and consumes these results in the reference's order.
// NOT COMPILED in the build image (no Go toolchain there).
package genomeGraph

/*
#cgo LDFLAGS: -lgnxalign
#include "gnxalign.h"
*/
import "C"

import (
	"log"
	"unsafe"

	"github.com/vertgenlab/gonomics/cigar"
	"github.com/vertgenlab/gonomics/dna"
	"github.com/vertgenlab/gonomics/fastq"
	"github.com/vertgenlab/gonomics/giraf"
)

func gnxCheck(ctx *C.gnx_ctx, rc C.int) {
	switch rc {
	case C.GNX_OK:
	case C.GNX_EBASE:
		panic("runtime error: index out of range")
	default:
		log.Panicf("gnxalign: %s", C.GoString(C.gnx_last_error(ctx)))
	}
}

func concatBases(seqs [][]dna.Base) ([]dna.Base, []int64) {
	off := make([]int64, len(seqs)+1)
	for i, s := range seqs {
		off[i+1] = off[i] + int64(len(s))
	}
	cat := make([]dna.Base, off[len(seqs)]+1)
	for i, s := range seqs {
		copy(cat[off[i]:], s)
	}
	return cat, off
}

// GpuSeedIndex replaces the map returned by IndexGenomeIntoMap (index.go:21-44) for edge-less nodes.
type GpuSeedIndex struct {
	ctx     *C.gnx_ctx
	h       *C.gnx_seed_index
	SeedLen int
}

func IndexGenomeIntoGpu(ctx unsafe.Pointer, genome []Node, seedLen int, seedStep int) *GpuSeedIndex {
	c := (*C.gnx_ctx)(ctx)
	seqs := make([][]dna.Base, len(genome))
	for i := range genome {
		if len(genome[i].Next) != 0 || len(genome[i].Prev) != 0 {
			log.Fatalf("Error: the GPU seed index covers linear references (nodes without edges)\n")
		}
		seqs[i] = genome[i].Seq
	}
	cat, off := concatBases(seqs)
	ix := &GpuSeedIndex{ctx: c, SeedLen: seedLen}
	gnxCheck(c, C.gnx_seed_index_new(c, (*C.uint8_t)(unsafe.Pointer(&cat[0])), (*C.int64_t)(unsafe.Pointer(&off[0])),
		C.int64_t(len(genome)), C.int(seedLen), C.int(seedStep), &ix.h))
	return ix
}

func (ix *GpuSeedIndex) Free() { C.gnx_seed_index_free(ix.h) }

// SeedMapBatch is seedMapMemPool for every read of a block: result[r] holds read r's seeds, already
// ordered with the reference's own final sort (search.go:596-600), so callers iterate exactly as before.
func (ix *GpuSeedIndex) SeedMapBatch(reads []fastq.FastqBig) [][]SeedDev {
	seqs := make([][]dna.Base, len(reads))
	for i := range reads {
		seqs[i] = reads[i].Seq
	}
	cat, off := concatBases(seqs)
	soff := make([]int64, len(reads)+1)
	seeds := make([]C.gnx_seed, 8*len(reads)+64)
	rc := C.gnx_seed_batch(ix.ctx, ix.h, (*C.uint8_t)(unsafe.Pointer(&cat[0])), (*C.int64_t)(unsafe.Pointer(&off[0])),
		C.int64_t(len(reads)), &seeds[0], (*C.int64_t)(unsafe.Pointer(&soff[0])), C.int64_t(len(seeds)))
	if rc == C.GNX_ECAP { // soff is filled: retry with the exact size
		seeds = make([]C.gnx_seed, soff[len(reads)]+1)
		rc = C.gnx_seed_batch(ix.ctx, ix.h, (*C.uint8_t)(unsafe.Pointer(&cat[0])), (*C.int64_t)(unsafe.Pointer(&off[0])),
			C.int64_t(len(reads)), &seeds[0], (*C.int64_t)(unsafe.Pointer(&soff[0])), C.int64_t(len(seeds)))
	}
	gnxCheck(ix.ctx, rc)
	out := make([][]SeedDev, len(reads))
	for r := range reads {
		lst := make([]SeedDev, 0, soff[r+1]-soff[r])
		for _, s := range seeds[soff[r]:soff[r+1]] {
			lst = append(lst, SeedDev{TargetId: uint32(s.target_id), TargetStart: uint32(s.target_start),
				QueryStart: uint32(s.query_start), Length: uint32(s.length), PosStrand: s.pos_strand != 0,
				TotalLength: uint32(s.total_length)})
		}
		if len(lst) > 100 { // the reference's own ordering, applied to the append-order list (search.go:596-600)
			SortSeedLen(lst)
		} else {
			heapSortSeeds(lst)
		}
		out[r] = lst
	}
	return out
}

// ExtendBatch runs LeftDynamicAln (left = true) or RightDynamicAln for every (alpha, beta) pair in one
// GPU call; routes are in traceback order, exactly as the reference returns them.
func ExtendBatch(ctx unsafe.Pointer, left bool, alphas, betas [][]dna.Base, scores [][]int64, gapPen int64) (score []int64, routes [][]cigar.Cigar, endI, endJ []int64) {
	c := (*C.gnx_ctx)(ctx)
	n := len(alphas)
	acat, aoff := concatBases(alphas)
	bcat, boff := concatBases(betas)
	dim := len(scores)
	flat := make([]int64, dim*dim)
	for i := range scores {
		copy(flat[i*dim:], scores[i])
	}
	side := C.int(C.GNX_EXT_RIGHT)
	if left {
		side = C.GNX_EXT_LEFT
	}
	score = make([]int64, n+1)
	endI, endJ = make([]int64, n+1), make([]int64, n+1)
	coff := make([]int64, n+1)
	cig := make([]cigar.Cigar, aoff[n]+boff[n]+int64(n)+1) // cigar.Cigar{RunLength int; Op byte} == gnx_cigar (16 B)
	p := func(s []int64) *C.int64_t { return (*C.int64_t)(unsafe.Pointer(&s[0])) }
	gnxCheck(c, C.gnx_extend_batch(c, side, (*C.uint8_t)(unsafe.Pointer(&acat[0])), p(aoff),
		(*C.uint8_t)(unsafe.Pointer(&bcat[0])), p(boff), C.int64_t(n), p(flat), C.int(dim), C.int64_t(gapPen), 1,
		p(score), p(endI), p(endJ), (*C.gnx_cigar)(unsafe.Pointer(&cig[0])), p(coff), C.int64_t(len(cig))))
	routes = make([][]cigar.Cigar, n)
	for i := 0; i < n; i++ {
		routes[i] = cig[coff[i]:coff[i+1]:coff[i+1]]
	}
	return score[:n], routes, endI[:n], endJ[:n]
}

// GswBatch is the whole per-read driver for a block of reads (gnx_gsw_batch): GraphSmithWatermanToGiraf
// (toGiraf.go:17-72) for every read -- WrapPairGiraf (:117-137) when paired, reads 2k / 2k+1 being the mates -- with
// the seed and extension steps of the block batched on the GPU and the reference's per-read loop replayed over their
// results inside the library.  The caller fills in what the aligner does not compute (QName, Seq, Qual, Notes) exactly
// as GraphSmithWatermanToGiraf does; RoutineFqPairToGiraf (routines.go) becomes "collect a block of pairs from the
// channel, call GswBatch, send the pairs on".
func (ix *GpuSeedIndex) GswBatch(reads []fastq.FastqBig, scores [][]int64, paired bool) []giraf.Giraf {
	n := len(reads)
	if n == 0 {
		return nil
	}
	seqs := make([][]dna.Base, n)
	for i := range reads {
		seqs[i] = reads[i].Seq
	}
	cat, off := concatBases(seqs)
	dim := len(scores)
	flat := make([]int64, dim*dim)
	for i := range scores {
		copy(flat[i*dim:], scores[i])
	}
	recs := make([]C.gnx_giraf, n)
	cig := make([]cigar.Cigar, 4*n+64) // cigar.Cigar{RunLength int; Op byte} == gnx_cigar (16 B)
	var need C.int64_t
	pairedFlag := C.int(0)
	if paired {
		pairedFlag = 1
	}
	call := func() C.int {
		return C.gnx_gsw_batch(ix.ctx, ix.h, (*C.uint8_t)(unsafe.Pointer(&cat[0])), (*C.int64_t)(unsafe.Pointer(&off[0])), C.int64_t(n),
			(*C.int64_t)(unsafe.Pointer(&flat[0])), C.int(dim), pairedFlag, &recs[0], (*C.gnx_cigar)(unsafe.Pointer(&cig[0])),
			C.int64_t(len(cig)), &need)
	}
	rc := call()
	if rc == C.GNX_ECAP && int(need) > len(cig) {
		cig = make([]cigar.Cigar, need)
		rc = call()
	}
	gnxCheck(ix.ctx, rc)
	out := make([]giraf.Giraf, n)
	for r := range reads {
		g := &recs[r]
		out[r] = giraf.Giraf{QName: reads[r].Name, QStart: int(g.q_start), QEnd: int(g.q_end), PosStrand: g.pos_strand != 0,
			Path: giraf.Path{TStart: int(g.t_start), TEnd: int(g.t_end)}, AlnScore: int(g.aln_score), MapQ: 255, Flag: uint8(g.flag),
			Seq: reads[r].Seq, Qual: reads[r].Qual, Notes: []giraf.Note{{Tag: []byte{'X', 'O'}, Type: 'Z', Value: "~"}}}
		if g.node >= 0 {
			out[r].Path.Nodes = []uint32{uint32(g.node)}
		}
		if g.n_cigar >= 0 {
			out[r].Cigar = cig[g.cigar_off : g.cigar_off+C.int64_t(g.n_cigar) : g.cigar_off+C.int64_t(g.n_cigar)]
		}
		if !out[r].PosStrand { // toGiraf.go:65-70
			out[r].Seq = reads[r].SeqRc
			fastq.ReverseQualUint8Record(out[r].Qual)
		}
	}
	return out
}
