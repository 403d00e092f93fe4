#!/usr/bin/env bash
# run_java_baseline.sh -- time the UNMODIFIED reference (Java) on the box's host cores, for a box that has
# JDK 8+ and Maven (or a prebuilt PLDA.jar).  SURVEY.md 8(d), BASELINE.md section 3.
#
# The build image of this repository has neither (no java / javac / mvn, no network), so bench.py's reference arm
# times the C port of the same sampler instead (oracle/lda_oracle.c oracle_baseline_sweeps, kind "port").  On a box
# with a JDK this script produces the number BASELINE.md's table leaves open: it runs the reference's own driver
# and reads the reference's own timers -- "Document sampling time" (z + count merge) and "Topic sampling time"
# (Phi), printed at the end of sample() (topics/UncollapsedParallelLDA.java:931-939).
#
#   scripts/run_java_baseline.sh <reference checkout> <corpus.txt> <scheme: ggs|pcgs|spalias|polyaurn> <topics> [iterations]
#
# The corpus file is in the reference's `name<TAB>label<TAB>text` format; write a synthetic one of a BASELINE shape with
#   python -c "import ldagroupedgibbssampler_b200 as L; L.write_synthetic_corpus('pubmed_shaped.txt', D=..., V=..., mean_len=...)"
# Settings that keep the comparison like-for-like:
#   start_diagnostic = 0, compute_likelihood = false   the O(D*K*V) diagnostic is excluded (UPL:707-853)
#   exec_time raised above the 10 s default            (configuration/LDAConfiguration.java:35, UPL:926-928)
#   topic_batches = #cores                              Phi on all cores (default is 2, LDAConfiguration.java:15)
#   heap sized for the K*V boxed AtomicIntegers         (UPL:102,363-368) and GGS's D x V thetaMatrix (GGS:36)
# To compare statistically with the GPU sampler pin the JVM to one core (taskset -c 0): the reference's
# docDirichletSampler / topicDirichletSampler fields race between workers.
set -euo pipefail
REF=${1:?reference checkout}; CORPUS=${2:?corpus file}; SCHEME=${3:?scheme}; TOPICS=${4:?topics}; ITERS=${5:-20}
command -v java >/dev/null || { echo "no JDK on this box: bench.py --impl reference times the C port instead" >&2; exit 3; }
CORES=$(nproc)
JAR=$(ls "$REF"/target/PLDA*.jar 2>/dev/null | head -1 || true)
if [ -z "$JAR" ]; then
  command -v mvn >/dev/null || { echo "no PLDA.jar under $REF/target and no mvn to build it" >&2; exit 3; }
  # pom.xml:152-154,269 hard-code /Users/clint/plda/target as the build directory: override it
  (cd "$REF" && mvn -q -DskipTests -Dproject.build.directory="$REF/target" package)
  JAR=$(ls "$REF"/target/PLDA*.jar | head -1)
fi
WORK=$(mktemp -d)
cat > "$WORK/run.cfg" <<CFG
configs = baseline
no_runs = 1
[baseline]
title = baseline
description = BASELINE.md section 3
dataset = $CORPUS
scheme = $SCHEME
seed = 2019
topics = $TOPICS
alpha = $(python3 -c "print(1.0 if $TOPICS <= 100 else 50.0 / $TOPICS)")
beta = 0.01
iterations = $ITERS
batches = $CORES
topic_batches = $CORES
rare_threshold = 0
topic_interval = 1000000
start_diagnostic = 0
compute_likelihood = false
exec_time = 86400
debug = 0
CFG
HEAP=${JAVA_HEAP:-64g}
( cd "$WORK" && java -Xmx"$HEAP" -jar "$JAR" --run_cfg="$WORK/run.cfg" ) | tee "$WORK/stdout.txt"
# the reference's own timers (ms, cumulative over the run)
Z_MS=$(grep -i "Document sampling time" "$WORK/stdout.txt" | grep -oE "[0-9]+" | tail -1)
PHI_MS=$(grep -i "Topic sampling time" "$WORK/stdout.txt" | grep -oE "[0-9]+" | tail -1)
TOKENS=$(awk -F'\t' '{n += split($3, a, " ")} END {print n}' "$CORPUS")
python3 - <<PY
z, phi, n, it, cores = float("${Z_MS:-0}"), float("${PHI_MS:-0}"), float("$TOKENS"), float("$ITERS"), $CORES
tot = (z + phi) / 1e3
print({"impl": "reference-java", "scheme": "$SCHEME", "topics": $TOPICS, "cores": cores, "iterations": it,
       "z_plus_merge_s": z / 1e3, "phi_s": phi / 1e3,
       "value": n * it / tot if tot > 0 else None, "unit": "tokens/s",
       "source": "Document sampling time / Topic sampling time, UncollapsedParallelLDA.java:931-939"})
PY
