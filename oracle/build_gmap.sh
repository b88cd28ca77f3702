#!/bin/bash
# oracle/build_gmap.sh -- ORACLE / INTEGRATION SCAFFOLDING (build container only; needs /root/reference).
#
# Builds two GMAP binaries from a scratch copy of the reference tree (nothing is written to /root/reference and no
# reference source is copied into this repository):
#   oracle/_ref/gmap_ref    the unmodified reference + oracle/genome_hr_standin.c (the checkout lacks genome_hr.c)
#   oracle/_ref/gmap_cuda   the same objects, except that the five gap-fill solvers of dynprog.c (and
#                           Dynprog_init/_setup/_term) have their definitions renamed *_cpu and are replaced by
#                           gmap-gsnap_b200/host/dynprog_dropin.c on top of libdynprog_cuda.so, and that gmap.c gets
#                           the two edits INTEGRATION.md describes (registering a user segment's genome blocks;
#                           worker threads created through Dynprog_cuda_worker_create, which runs DPC_FIBERS copies
#                           of worker_thread per OS thread so that stage 3's gaps reach the device in batches).
# It also stages the reference's own align.test inputs and golden output next to the binaries (git-ignored, but
# they travel to the GPU box) so that tests/test_gpu_gmap.py can run BASELINE config 1 there.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF:-/root/reference}"
WORK="${WORK:-/tmp/gmap_build}"
OUT="$REPO/oracle/_ref"
[ -d "$REF/src" ] || { echo "build_gmap.sh: $REF absent; keeping prebuilt binaries"; exit 0; }
mkdir -p "$OUT"
if [ ! -f "$WORK/src/Makefile" ]; then
  rm -rf "$WORK"; cp -r "$REF" "$WORK"; chmod -R u+w "$WORK"
  (cd "$WORK" && ./configure > configure.log 2>&1)
fi
cd "$WORK/src"
cp "$REPO/oracle/genome_hr_standin.c" genome_hr.c
make -j"$(nproc)" gmap > make.log 2>&1
cp gmap "$OUT/gmap_ref"

CC="${CC:-gcc}"
CFLAGS=(-DHAVE_CONFIG_H -I. -mpopcnt '-DTARGET="x86_64-unknown-linux-gnu"' '-DGMAPDB="/usr/share/gmapdb"' -O3)
SYMS="Dynprog_init Dynprog_setup Dynprog_term Dynprog_single_gap Dynprog_cdna_gap Dynprog_genome_gap Dynprog_end5_gap Dynprog_end3_gap Dynprog_end5_splicejunction Dynprog_end3_splicejunction Dynprog_microexon_int"
# Only the DEFINITIONS of the replaced functions are renamed (<name>_cpu), in a scratch copy of dynprog.c: the
# reference writes a function's name at the start of a line where it defines it and indented where it calls it.
# Calls from the unreplaced functions of dynprog.c (Dynprog_end5_known / Dynprog_end3_known call Dynprog_end5_gap /
# Dynprog_end3_gap, dynprog.c:6474-6900) therefore reach the drop-in too; nothing in the product calls a _cpu solver
# (Dynprog_init/_setup/_term_cpu are called by the drop-in to keep the reference's own tables for the functions
# that are not replaced).  A maintainer would simply delete the replaced bodies from dynprog.c.
SEDARGS=(); for s in $SYMS; do SEDARGS+=(-e "s/^$s (/${s}_cpu (/"); done
sed "${SEDARGS[@]}" dynprog.c > cuda-dynprog_cpu.c
for s in $SYMS; do grep -q "^${s}_cpu (" cuda-dynprog_cpu.c; done
$CC "${CFLAGS[@]}" -w -ffunction-sections -c cuda-dynprog_cpu.c -o cuda-dynprog_cpu.o
$CC "${CFLAGS[@]}" -I"$REPO/include" -Wall -c "$REPO/gmap-gsnap_b200/host/dynprog_dropin.c" -o cuda-dynprog_dropin.o
# the one-line change to gmap.c: hand the user segment's blocks to the library
sed 's|^    Genome_user_setup(genome_blocks);|    Genome_user_setup(genome_blocks);\n    { extern void Dynprog_cuda_register_blocks (UINT4 *blocks, unsigned int nwords); Dynprog_cuda_register_blocks(genome_blocks,((Sequence_fulllength(usersegment) + 31)/32U)*3 + 4); }|' gmap.c > cuda-gmap.c
grep -q Dynprog_cuda_register_blocks cuda-gmap.c
# the second change: workers are created through the drop-in's fiber scheduler
sed -i 's|pthread_create(&(worker_thread_ids\[i\]),&thread_attr_join,worker_thread,(void \*) NULL);|{ extern int Dynprog_cuda_worker_create (pthread_t *, const pthread_attr_t *, void *(*) (void *), void *); Dynprog_cuda_worker_create(\&(worker_thread_ids[i]),\&thread_attr_join,worker_thread,(void *) NULL); }|' cuda-gmap.c
grep -q Dynprog_cuda_worker_create cuda-gmap.c
$CC "${CFLAGS[@]}" -DExcept_stack_create=Dynprog_cuda_except_stack_create -DExcept_stack_destroy=Dynprog_cuda_except_stack_destroy \
    -c cuda-gmap.c -o cuda-gmap.o
OBJS=$(ls gmap-*.o | grep -v -e '^gmap-dynprog.o$' -e '^gmap-gmap.o$')
$CC -O3 -o "$OUT/gmap_cuda" $OBJS cuda-dynprog_cpu.o cuda-dynprog_dropin.o cuda-gmap.o \
    -L"$REPO/gmap-gsnap_b200/csrc" -ldynprog_cuda -Wl,-rpath,'$ORIGIN/../../gmap-gsnap_b200/csrc' -Wl,--gc-sections -lz -lm -lpthread
# link-time proof that no CPU solver is reachable: every function of the renamed dynprog.c sits in its own section and
# the linker drops what nothing references -- the replaced solver bodies must be gone from the binary
for s in $SYMS; do
  case $s in Dynprog_init|Dynprog_setup|Dynprog_term) continue;; esac
  if nm "$OUT/gmap_cuda" | grep -q " T ${s}_cpu$"; then echo "build_gmap.sh: ${s}_cpu is still reachable in gmap_cuda" >&2; exit 1; fi
done
nm "$OUT/gmap_cuda" | grep -c "_cpu$" > /dev/null
# index-building tools of the reference (for the whole-program bench on a synthetic genome database, BASELINE config 5)
make -j"$(nproc)" gmapindex iit_store >> make.log 2>&1
(cd ../util && make -s fa_coords gmap_process gmap_build >> ../src/make.log 2>&1)
mkdir -p "$OUT/bin"
cp gmapindex iit_store ../util/fa_coords ../util/gmap_process ../util/gmap_build "$OUT/bin/"
# measurement scaffolding: the unmodified reference with a timer around its five gap-fill solvers
ARGS5=""; for s in Dynprog_single_gap Dynprog_cdna_gap Dynprog_genome_gap Dynprog_end5_gap Dynprog_end3_gap; do ARGS5="$ARGS5 --redefine-sym $s=${s}_cpu"; done
objcopy $ARGS5 gmap-dynprog.o timed-dynprog_cpu.o
$CC "${CFLAGS[@]}" -c "$REPO/oracle/dynprog_timed.c" -o timed-wrap.o
$CC -O3 -o "$OUT/gmap_ref_timed" $(ls gmap-*.o | grep -v -e '^gmap-dynprog.o$') timed-dynprog_cpu.o timed-wrap.o -lz -lm -lpthread
mkdir -p "$OUT/align_test"
cp "$REF/tests/ss.her2" "$REF/tests/ss.chr17test" "$REF/tests/align.test.ok" "$OUT/align_test/"
# sanity: the unmodified build reproduces the reference's golden output
"$OUT/gmap_ref" -A -g "$OUT/align_test/ss.chr17test" "$OUT/align_test/ss.her2" 2>/dev/null | diff -q - "$OUT/align_test/align.test.ok"
echo "build_gmap.sh: gmap_ref (align.test identical) and gmap_cuda built in $OUT"
