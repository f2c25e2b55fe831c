//go:build cuda

package align

/*
#include <stdlib.h>
#include "gnxalign.h"
*/
import "C"

import (
	"log"
	"math"
	"runtime"
	"sync"
	"unsafe"

	"github.com/vertgenlab/gonomics/dna"
	"github.com/vertgenlab/gonomics/dna/dnaTwoBit"
	"github.com/vertgenlab/gonomics/fasta"
)

// ---- several GPUs behind one call (gnx_multi_*) ---------------------------------------------------------------
// One gnx_multi for the process: a context per visible device, batches cut into contiguous cell-balanced shards by
// the library, results written into the caller's slices in pair order.  NOT COMPILED in the build image.
var (
	multiOnce sync.Once
	multi     *C.gnx_multi
	multiMu   sync.Mutex // a gnx_multi runs one batch at a time
)

func getMulti() *C.gnx_multi {
	multiOnce.Do(func() {
		multi = C.gnx_multi_create(nil, 0, C.size_t(workspaceBytes)) // every visible device
		if multi == nil {
			log.Panicf("gnxalign: %s", C.GoString(C.gnx_multi_last_error(nil)))
		}
	})
	return multi
}

// MultiAffineGapBatch is AffineGapBatchCat over every GPU of the box (gnx_multi_affine_batch).
func MultiAffineGapBatch(acat []dna.Base, aoff []int64, bcat []dna.Base, boff []int64, scores [][]int64, gapOpen, gapExtend int64, mode int) ([]int64, []int64, []Cigar) {
	n := len(aoff) - 1
	if n <= 0 {
		return nil, []int64{0}, nil
	}
	mg := getMulti()
	multiMu.Lock()
	defer multiMu.Unlock()
	flat, dim := flatten(scores)
	out := make([]int64, n)
	coff := make([]int64, n+1)
	cig := make([]Cigar, 16*n+64)
	rc := C.gnx_multi_affine_batch(mg, basePtr(acat), i64Ptr(aoff), basePtr(bcat), i64Ptr(boff), C.int64_t(n), i64Ptr(flat), C.int(dim),
		C.int64_t(gapOpen), C.int64_t(gapExtend), C.int(mode), 1, i64Ptr(out), cigPtr(cig), i64Ptr(coff), C.int64_t(len(cig)))
	if rc == C.GNX_ECAP {
		cig = make([]Cigar, coff[n])
		rc = C.gnx_multi_copy_last_cigars(mg, cigPtr(cig), C.int64_t(len(cig)))
	}
	runtime.KeepAlive(flat)
	switch rc {
	case C.GNX_OK:
	case C.GNX_EBASE:
		panic("runtime error: index out of range (base >= len(scores))")
	default:
		log.Panicf("gnxalign: %s", C.GoString(C.gnx_multi_last_error(mg)))
	}
	return out, coff, cig[:coff[n]]
}

// PackTwoBit concatenates dnaTwoBit.TwoBit sequences into the (words, lens) form AffineGapBatchTwoBit takes.
func PackTwoBit(seqs []dnaTwoBit.TwoBit) ([]uint64, []int64) {
	n := 0
	for i := range seqs {
		n += (seqs[i].Len + 31) / 32
	}
	words := make([]uint64, 0, n)
	lens := make([]int64, len(seqs))
	for i := range seqs {
		words = append(words, seqs[i].Seq[:(seqs[i].Len+31)/32]...)
		lens[i] = int64(seqs[i].Len)
	}
	return words, lens
}

// PackBasesUniform is dnaTwoBit.NewTwoBit of count sequences of length n held back to back as dna.Base bytes, done by
// the library's host threads (gnx_pack_twobit_host: the packer gnx_affine_batch itself runs while it stages a large
// uniform batch, so callers of AffineGapBatchCat get it implicitly).  Panics like scores[a][b] for a base >= 4.
func PackBasesUniform(cat []dna.Base, count int, n int) []uint64 {
	words := make([]uint64, count*((n+31)/32))
	if count == 0 || n == 0 {
		return words
	}
	rc := C.gnx_pack_twobit_host((*C.uint8_t)(unsafe.Pointer(&cat[0])), C.int64_t(count), C.int64_t(n), (*C.uint64_t)(unsafe.Pointer(&words[0])))
	if rc == C.GNX_EBASE {
		panic("runtime error: index out of range")
	} else if rc != C.GNX_OK {
		log.Panicf("gnxalign: gnx_pack_twobit_host failed (%d)", int(rc))
	}
	return words
}

// ---- the profile DP and the progressive multiple alignment (cmd/faChunkAlign) -----------------------------------
// multipleAffineGap / multipleAffineGapChunk (align/affineGap_highMem.go:274-353) keep their signatures; nearestGroups /
// nearestGroupsChunk (align/multiAlign.go:27-57) evaluate ALL group pairs of a round in one GPU call
// (gnx_multi_affine_chunk_batch), with the reference's "first strictly larger score wins" scan over the results.

func stackGroups(groups [][]fasta.Fasta) ([]dna.Base, []int64, []int64) {
	off := make([]int64, len(groups)+1)
	nseq := make([]int64, len(groups))
	for g := range groups {
		nseq[g] = int64(len(groups[g]))
		var b int64
		for _, f := range groups[g] {
			b += int64(len(f.Seq))
		}
		off[g+1] = off[g] + b
	}
	cat := make([]dna.Base, off[len(groups)])
	p := int64(0)
	for g := range groups {
		for _, f := range groups[g] {
			copy(cat[p:], f.Seq)
			p += int64(len(f.Seq))
		}
	}
	return cat, off, nseq
}

func profileBatch(groups [][]fasta.Fasta, px, py []int64, scores [][]int64, gapOpen, gapExtend, chunkSize int64) ([]int64, []int64, []Cigar) {
	ctx := getCtx()
	defer putCtx(ctx)
	flat, dim := flatten(scores)
	cat, off, nseq := stackGroups(groups)
	n := len(px)
	out := make([]int64, n)
	coff := make([]int64, n+1)
	cig := make([]Cigar, 64*n+64)
	rc := C.gnx_multi_affine_chunk_batch(ctx, basePtr(cat), i64Ptr(off), i64Ptr(nseq), C.int64_t(len(groups)), i64Ptr(px), i64Ptr(py),
		C.int64_t(n), i64Ptr(flat), C.int(dim), C.int64_t(gapOpen), C.int64_t(gapExtend), C.int64_t(chunkSize), 1,
		i64Ptr(out), cigPtr(cig), i64Ptr(coff), C.int64_t(len(cig)))
	if rc == C.GNX_ECAP {
		cig = make([]Cigar, coff[n])
		rc = C.gnx_copy_last_cigars(ctx, cigPtr(cig), C.int64_t(len(cig)))
	}
	runtime.KeepAlive(flat)
	check(ctx, rc)
	return out, coff, cig
}

func multipleAffineGapChunk(alpha []fasta.Fasta, beta []fasta.Fasta, scores [][]int64, gapOpen int64, gapExtend int64, chunkSize int64) (int64, []Cigar) {
	sc, coff, cig := profileBatch([][]fasta.Fasta{alpha, beta}, []int64{0}, []int64{1}, scores, gapOpen, gapExtend, chunkSize)
	return sc[0], cig[coff[0]:coff[1]]
}

func multipleAffineGap(alpha []fasta.Fasta, beta []fasta.Fasta, scores [][]int64, gapOpen int64, gapExtend int64) (int64, []Cigar) {
	return multipleAffineGapChunk(alpha, beta, scores, gapOpen, gapExtend, 1)
}

func nearestGroupsChunk(groups [][]fasta.Fasta, scoreMatrix [][]int64, gapOpen int64, gapExtend int64, chunkSize int) (bestX int, bestY int, bestScore int64, bestRoute []Cigar) {
	var px, py []int64
	for x := 0; x < len(groups)-1; x++ {
		for y := x + 1; y < len(groups); y++ {
			px, py = append(px, int64(x)), append(py, int64(y))
		}
	}
	bestScore = math.MinInt64
	if len(px) == 0 {
		return
	}
	sc, coff, cig := profileBatch(groups, px, py, scoreMatrix, gapOpen, gapExtend, int64(chunkSize))
	for p := range px { // the reference's scan order (x outer, y inner), strict ">"
		if sc[p] > bestScore {
			bestX, bestY, bestScore, bestRoute = int(px[p]), int(py[p]), sc[p], cig[coff[p]:coff[p+1]:coff[p+1]]
		}
	}
	return
}

func nearestGroups(groups [][]fasta.Fasta, scoreMatrix [][]int64, gapOpen int64, gapExtend int64) (int, int, int64, []Cigar) {
	return nearestGroupsChunk(groups, scoreMatrix, gapOpen, gapExtend, 1)
}

var _ = unsafe.Pointer(nil)
