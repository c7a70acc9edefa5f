// The transposed CTA-pair tokeniser with 96 rows per CTA (192-row pair tiles): same source as rq_tcx.cu (64 rows), see the tile
// shape notes there.  rqb200_tokenize_tc_run picks the shape by batch size (csrc/rq_tc.cu).
#define TX_R 96
#include "rq_tcx.cu"
