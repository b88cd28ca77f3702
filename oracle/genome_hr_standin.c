/* oracle/genome_hr_standin.c -- ORACLE SCAFFOLDING, NOT PRODUCT CODE.
 *
 * The checkout under /root/reference lacks src/genome_hr.c (/root/reference/.MISSING_LARGE_BLOBS), so gmap does
 * not link.  This file supplies, for the oracle build of gmap only (oracle/build_gmap.sh), the ten symbols the
 * gmap link needs.  Their behaviour is specified inside the reference itself:
 *   - Genome_prev_{donor,acceptor,antidonor,antiacceptor}_position must equal the tables computed by
 *     find_canonical_dinucleotides (stage2.c:742-863); check_canonical_dinucleotides_hr (stage2.c:900-971)
 *     asserts exactly that equivalence:
 *       donor / antiacceptor: largest p <= pos with seg[p+1..p+2] == "GT" / "CT"
 *       acceptor / antidonor: largest q <= pos with seg[q-2..q-1] == "AG" / "AC"
 *       -1 when there is none; seg = genome[genomicstart, genomicend) or its reverse complement when !plusp.
 *   - the gamma decoders (index databases, `gmap -d`) invert write_gamma / Indexdb_write_gammaptrs
 *     (indexdb.c:1212-1252, 1978-2037); the whole-program bench builds its database with k-mer = base size 12
 *     like tests/setup1.test.in:12, where every block is a single absolute offset.
 * It reads bases through the same 3-word blocks as genome.c:9325-9362.
 */
#include <stdio.h>
#include <stdlib.h>
#include "bool.h"
#include "types.h"
#include "genomicpos.h"
#include "mode.h"
#include "genome_hr.h"

static UINT4 *ref_blocks;

void
Genome_hr_setup (UINT4 *ref_blocks_in, UINT4 *snp_blocks_in,
		 bool query_unk_mismatch_p_in, bool genome_unk_mismatch_p_in, Mode_T mode_in) {
  (void) snp_blocks_in; (void) query_unk_mismatch_p_in; (void) genome_unk_mismatch_p_in; (void) mode_in;
  ref_blocks = ref_blocks_in;
}

void
Genome_hr_user_setup (UINT4 *ref_blocks_in,
		      bool query_unk_mismatch_p_in, bool genome_unk_mismatch_p_in, Mode_T mode_in) {
  (void) query_unk_mismatch_p_in; (void) genome_unk_mismatch_p_in; (void) mode_in;
  ref_blocks = ref_blocks_in;
}

/* Elias-gamma offsets: the inverse of write_gamma / Indexdb_write_gammaptrs (indexdb.c:1212-1252, 1978-2037), used
   the way Indexdb_offsets_from_gammas uses it (indexdb.c:1395-1416): a block starts with an absolute offset word,
   followed by blocksize-1 gamma codes of (difference + 1), most significant bit first in 32-bit words; ctr is the
   number of unread bits left in *ptr (0 = *ptr is the next unread word). */
static unsigned int
gamma_bit (unsigned int **ptr, int *ctr) {
  unsigned int bit;
  if (*ctr == 0) *ctr = 32;
  bit = (**ptr >> (*ctr - 1)) & 1U;
  if (--*ctr == 0) (*ptr)++;
  return bit;
}

int
Genome_read_gamma (unsigned int **ptr, int ctr, unsigned int *cum) {
  unsigned int value = 1U;
  int nzeros = 0;
  while (gamma_bit(ptr,&ctr) == 0U) nzeros++;
  while (nzeros-- > 0) value = (value << 1) | gamma_bit(ptr,&ctr);
  *cum += value - 1U;
  return ctr;
}

Positionsptr_T
Genome_offsetptr_from_gammas (Positionsptr_T *end0, UINT4 *gammaptrs, Positionsptr_T *offsetscomp,
			      unsigned int offsets_blocksize, Storedoligomer_T oligo) {
  unsigned int block = oligo/offsets_blocksize, rem = oligo % offsets_blocksize, j;
  unsigned int *ptr = &offsetscomp[gammaptrs[block]];
  unsigned int cum = *ptr++;
  Positionsptr_T ptr0;
  int ctr = 0;
  for (j = 0; j < rem; j++) ctr = Genome_read_gamma(&ptr,ctr,&cum);
  ptr0 = cum;
  if (rem + 1U == offsets_blocksize) {
    *end0 = offsetscomp[gammaptrs[block+1]];
  } else {
    ctr = Genome_read_gamma(&ptr,ctr,&cum);
    *end0 = cum;
  }
  return ptr0;
}

Positionsptr_T
Genome_offsetptr_only_from_gammas (UINT4 *gammaptrs, Positionsptr_T *offsetscomp,
				   unsigned int offsets_blocksize, Storedoligomer_T oligo) {
  Positionsptr_T end0;
  return Genome_offsetptr_from_gammas(&end0,gammaptrs,offsetscomp,offsets_blocksize,oligo);
}

static char
base_at (Genomicpos_T pos) {
  UINT4 *b = ref_blocks + (size_t) (pos/32U)*3;
  int bit = (int) (pos % 32U);
  if (b[2] & (1U << bit)) return 'N';
  return "ACGT"[(bit < 16 ? b[1] >> (2*bit) : b[0] >> (2*bit - 32)) & 3U];
}

/* seg[i] of the segment as stage 2 sees it */
static char
seg_at (int i, Genomicpos_T genomicstart, Genomicpos_T genomicend, bool plusp) {
  char c;
  if (i < 0 || (Genomicpos_T) i >= genomicend - genomicstart) return 'N';
  if (plusp == true) return base_at(genomicstart + (Genomicpos_T) i);
  c = base_at(genomicend - 1U - (Genomicpos_T) i);
  return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}

static int
prev_left (int pos, Genomicpos_T genomicstart, Genomicpos_T genomicend, bool plusp, char c1, char c2) {
  int p;
  for (p = pos; p >= 1; p--) {
    if (seg_at(p+1,genomicstart,genomicend,plusp) == c1 && seg_at(p+2,genomicstart,genomicend,plusp) == c2) return p;
  }
  return -1;
}

static int
prev_right (int pos, Genomicpos_T genomicstart, Genomicpos_T genomicend, bool plusp, char c1, char c2) {
  int q;
  for (q = pos; q >= 4; q--) {
    if (seg_at(q-2,genomicstart,genomicend,plusp) == c1 && seg_at(q-1,genomicstart,genomicend,plusp) == c2) return q;
  }
  return -1;
}

int
Genome_prev_donor_position (int pos, Genomicpos_T genomicstart, Genomicpos_T genomicend, int pos5, bool plusp) {
  (void) pos5; return prev_left(pos,genomicstart,genomicend,plusp,'G','T');
}
int
Genome_prev_antiacceptor_position (int pos, Genomicpos_T genomicstart, Genomicpos_T genomicend, int pos5, bool plusp) {
  (void) pos5; return prev_left(pos,genomicstart,genomicend,plusp,'C','T');
}
int
Genome_prev_acceptor_position (int pos, Genomicpos_T genomicstart, Genomicpos_T genomicend, int pos5, bool plusp) {
  (void) pos5; return prev_right(pos,genomicstart,genomicend,plusp,'A','G');
}
int
Genome_prev_antidonor_position (int pos, Genomicpos_T genomicstart, Genomicpos_T genomicend, int pos5, bool plusp) {
  (void) pos5; return prev_right(pos,genomicstart,genomicend,plusp,'A','C');
}
