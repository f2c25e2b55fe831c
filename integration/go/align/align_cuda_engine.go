//go:build cuda

package align

import "github.com/vertgenlab/gonomics/dna"

// GoAffineGapLocalEngine keeps the reference's channel interface (align/affineGap_highMem.go:120-179:
// buffered channels of capacity 1000, one goroutine, FIFO output, outputs closed when inputs close) but
// drains whatever is queued into ONE GPU batch per iteration, so a producer that keeps the channel full
// gets batch-sized launches without changing a line.  NOT COMPILED in the build image (no Go toolchain).
func GoAffineGapLocalEngine(scores [][]int64, gapOpen int64, gapExtend int64) (inputs chan<- TargetQueryPair, outputs <-chan TargetQueryPair) {
	i := make(chan TargetQueryPair, 1000)
	o := make(chan TargetQueryPair, 1000)
	go func() {
		const maxBatch = 1 << 18 // one kernel chunk of the library (gnx_set_option "chunk_pairs")
		batch := make([]TargetQueryPair, 0, maxBatch)
		for first := range i {
			batch = append(batch[:0], first)
		drain:
			for len(batch) < maxBatch {
				select {
				case p, ok := <-i:
					if !ok {
						break drain
					}
					batch = append(batch, p)
				default:
					break drain
				}
			}
			targets := make([][]dna.Base, len(batch))
			queries := make([][]dna.Base, len(batch))
			for k := range batch {
				targets[k], queries[k] = batch[k].Target, batch[k].Query
			}
			sc, routes := AffineGapBatch(targets, queries, scores, gapOpen, gapExtend, 1)
			for k := range batch {
				batch[k].Score, batch[k].Cigar = sc[k], routes[k]
				o <- batch[k]
			}
		}
		close(o)
	}()
	return i, o
}
