mkdir -p gpurun_out
# memcheck of what round 2 added to the design loop (alternative structures + snake moves, motifs, pseudoknot overlay, negative design),
# of the restructured generic (two-strand) kernels, the 2-best DP and the wide exterior kernels
cat > /tmp/san2.py <<'P'
import os, sys, random
sys.path.insert(0, os.getcwd())
import numpy as np
from desirna_b200 import design, engine
from desirna_b200.utils import stats_inputs_outputs as sio
engine.init(0); engine.params_builtin(1999)
alt = sio.make_input("alt", "((((((....))))))....((((....))))")
alt.add_alt_sec_struct(["....((((((....))))))((((....))))", "((((((....))))))....((((....))))"])
random.seed(0)
loop = design.DesignLoop([alt, sio.make_input("p", "((((....))))")], design.DesignOptions(replicas=4, RE_attempt=6, motifs={"GNRA": -1.0}), seed=1)
loop.run(2); print("alt", loop.jobs()["solved_step"]); loop.close()
loop = design.DesignLoop([sio.make_input("pk", "((((((....[[[[..))))))......]]]]....")], design.DesignOptions(replicas=4, RE_attempt=6, pks="on"), seed=2)
loop.run(2); print("pk", loop.jobs()["mfe_ss"]); loop.close()
loop = design.DesignLoop([sio.make_input("nd", "((((....))))")], design.DesignOptions(replicas=4, RE_attempt=10, subopt="on"), seed=3)
loop.run(3); print("nd", loop.replicas()["rec"][:, 14].tolist()); loop.close()
het = sio.make_input("het", "(((.(((((....))..&(((....)))..))))))", "NNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNNN")
loop = design.DesignLoop([het], design.DesignOptions(replicas=4, RE_attempt=6, oligo_state="heterodimer"), seed=4)
loop.run(2); print("het", loop.jobs()["mfe_ss"]); loop.close()
rng = np.random.default_rng(1)
seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, 30)) + "&" + "".join("ACGU"[x] for x in rng.integers(0, 4, 25)) for _ in range(4)]
print("two-strand", engine.score_batch(seqs, want=7)["mfe_dcal"].tolist())
seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, n)) for n in (130, 77, 20)]
print("wide ext", engine.score_batch(seqs, want=7)["mfe_dcal"].tolist())
print("second best", [x.tolist() for x in engine.second_best(seqs)])
P
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/r02_sanitize_memcheck_design.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r02_sanitize_memcheck_design.log
