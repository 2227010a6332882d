"""Regenerate INTEGRATION.md: prose below + the integration/ sources embedded verbatim (tests/test_host_cpu.py keeps the
two identical)."""
import os

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def src(rel):
    return open(os.path.join(ROOT, rel)).read()


MD = r'''# INTEGRATION — wiring `libflashfry_b200.so` into FlashFry (Scala/JVM)

The reference has no FFI today.  The two seams the library replaces are Scala traits, so the integration is: one Java
class with `native` methods, one JNI shim (C), two small Scala classes.  This repository's build image has no JDK, so the
Java and Scala sources are shipped uncompiled for a FlashFry maintainer; the **C side is executed**: `tests/stubs/jni.h`
+ `tests/stubs/jni_mock.c` are a functional stand-in for the JNIEnv functions the shim uses (heap-backed arrays, strings,
direct buffers), and `tests/stubs/jni_exec.c` calls every `Java_flashfry_NativeBridge_*` entry point through it against
the real library on the GPU and compares with direct calls of the C ABI (`tests/test_gpu_jni.py`).  The C ABI itself is
what all other tests exercise (ctypes, and the C++ host mirror `flashfry_b200/csrc/host/`).

## 1. What replaces what

| FlashFry call site | today | with the library |
|---|---|---|
| `modules/OffTargetDiscovery.scala:109-110` | `new OrderedBinTraversalFactory(...)`: 4^7 bins x G guides tested up front (1.6e9 JVM compares for 100 000 guides) | **skipped with `--gpu`**: the native side prunes by itself; pass `LinearTraversal` (or `null`) |
| `modules/OffTargetDiscovery.scala:119-135` | `LinearTraverser.scan(...)` / `SeekTraverser.scan(...)` | `new GpuTraverser(maximumOffTargets, positionOutput).scan(...)` → `ff_load_database` + `ff_discover` |
| `modules/OffTargetDiscovery.scala:141-152` + `targetio/TabDelimitedHandler.scala:132-154` | one `CRISPRHit` per hit (1e7 objects for 100 000 guides), then `TabDelimitedOutput.write` | `GpuTraverser.scanToTsv` → `ff_hits_write_tsv`: the TSV straight from the CSR (byte-identical: `EMX1.output` md5 through the CLI) |
| `modules/ScoreResults.scala:169-183` (`"hsu2013"`, `"doench2016cfd"`) | `new CrisprMitEduOffTarget()`, `new Doench2016CFDScore()` | `new GpuScoreModel(FF_METRIC_…, ctx)` → `ff_score_enzyme` |
| (optional) discover + score in one go | two CLI runs through a TSV | `ff_discover_score`: the hit list is scored while still in HBM |
| `scoring/ClosestHit.scala:43-76` (`minot`), `DangerousSequences.scala:61-65` (in-genome count) | per-guide loops over the hit list | `ff_hit_aggregates` (integer reductions on the GPU) |
| (optional) cold start | BGZF inflate + block walk on every run (~40 s on hg38) | `ff_load_database`: 2.8 s for a 3e8-target database on 16 host threads; `ff_save_image` / `ff_load_image`: a flat side-car (not a FlashFry format) |
| (optional) every GPU of the box | the JVM is one process | `multiCreate` / `multiDiscover` → `ff_multi_*`: guide shards on every device, one NCCL all-gather of the totals; `multiSetOption("shard_mode", 1)` shards the INDEX WORK instead (every device scans 1/n of the index for all guides, candidates travel to the guide's owner over NVLink peer memory; `ff_shard.inl`) -- same rows, faster from 4 GPUs on |
| (optional) fewer bytes over PCIe | — | `setOption("compact_hits", 1)`: hit lists carry 32-bit database indices; `dbHostTargets` is the target array as a direct buffer |

An extension outside FlashFry's feature set, `ff_discover_bulge` (≤ k mismatches plus one 1-bp RNA or DNA bulge, defined
in `include/flashfry_b200.h`), binds exactly like `ff_discover` with one more `int bulgeFlags` argument and one more
`byte[] bulge` result array; a `--bulge rna,dna` option on `discover` selects it (implemented in
`flashfry_b200_cli`; bulged tokens are written `SEQ_count_mm_R<q>` / `SEQ_count_mm_D<q>`).  It has no counterpart in
the reference, so it is not part of the drop-in contract.

The sources below are also shipped as files — `integration/java/flashfry/NativeBridge.java`,
`integration/jni/flashfry_b200_jni.c`, `integration/scala/GpuTraverser.scala`, `integration/scala/GpuScoreModel.scala` —
and `tests/test_host_cpu.py` keeps them identical to this document, compiles the shim with `-Wall -Wextra -Werror`
against the stand-in `<jni.h>` and checks that every JNI entry point has its `native` declaration.

## 2. Java native declarations (`src/main/java/flashfry/NativeBridge.java`)

```java
%(java)s```

## 3. JNI shim (`flashfry_b200_jni.c`)

JNI rules the shim keeps: no `GetPrimitiveArrayCritical` (a critical region around a GPU call would stall every JVM
thread that needs a GC for the whole discover) — arrays are copied with `Get<T>ArrayRegion` into `malloc`'d buffers;
array lengths are validated before the native side reads them (`rowPtr.length == nGuides + 1`, score arrays, the TSV
guide descriptions); every failure becomes an `IllegalStateException` carrying `ff_last_error()`.

```c
%(c)s```

Build: `gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude flashfry_b200_jni.c -Lflashfry_b200 -lflashfry_b200 -o libflashfry_b200_jni.so`.

## 4. Scala glue

```scala
%(scala1)s```

In `modules/OffTargetDiscovery.scala` add a `--gpu` option (or honour `FLASHFRY_GPU`): **do not construct
`OrderedBinTraversalFactory`** (`:109-110`; its precompute alone is 1.6e9 comparisons for 100 000 guides), build
`new GpuTraverser(maximumOffTargets, positionOutput)` (both values are already parsed there, `:57-59`, `:50`) and call
`scan` with the same arguments instead of the two existing cases — or `scanToTsv` when the run ends in
`TabDelimitedOutput` anyway; set the overflow callback to a no-op (`guideStorage.setTraversalOverFlowCallback(_ => ())`)
because no traversal needs to be told.  `CRISPRSiteOT`'s overflow budget is a constructor argument without an accessor
(`crispr/CRISPRSiteOT.scala:31`), which is why the traverser takes it as its own argument.

```scala
%(scala2)s```

## 5. Python (ctypes) — what the tests use

`flashfry_b200/_native.py` declares every symbol of the header; `flashfry_b200/api.py` wraps them (`Context.load_database`,
`.discover`, `.score`, `.discover_score`, `.discover_device`, `.timings`, `.set_option`; `MultiContext`).  The C++ mirror
(`csrc/host/flashfry_host.hpp`: `GpuTraverser::scan`, `GpuScoreModel::scoreGuides`) is the same call sequence as the Scala
glue above; `flashfry_b200_cli discover` takes the TSV fast path (`ff_hits_write_tsv`), `host_selftest traverse` the
object-building one, and `tests/test_gpu_jni.py` checks that the two files are identical.

## 6. Error mapping

| library | reference behaviour it stands for |
|---|---|
| `FF_EFORMAT` "doesn't have the magic number…" / "…correct version" | asserts in `BinaryHeader.scala:121-124` |
| `FF_EFORMAT` "Invalid bin type…" | `IllegalStateException` in `BlockManager.scala:85-87` |
| `FF_EFORMAT` "Failed to correctly parse block…" | `require` in `BlockManager.scala:236-237` |
| `FF_EFORMAT` truncated BGZF member / implausible image sizes | htsjdk would throw while seeking |
| `FF_EINVAL` "Unable to find the correct parameter pack…" | `StandardScanParameters.scala:69` |
| `FF_EINVAL` malformed CSR handed to `ff_score` / `ff_hit_aggregates` | (no counterpart: the JVM owns its lists) |
| `FF_EUNSUPPORTED` from `ff_score` on a non-Cas9-23 pack | host prints `NA` (`ScoreModel.scala:125-128`) |
| `FF_ENOMEM` / `FF_EIO` "internal error" | a C++ exception never crosses the ABI (every entry point is guarded) |
| `FF_ENODEVICE` | no equivalent: there is deliberately no CPU fallback |
'''

if __name__ == "__main__":
    out = MD % {"java": src("integration/java/flashfry/NativeBridge.java"), "c": src("integration/jni/flashfry_b200_jni.c"),
                "scala1": src("integration/scala/GpuTraverser.scala"), "scala2": src("integration/scala/GpuScoreModel.scala")}
    open(os.path.join(ROOT, "INTEGRATION.md"), "w").write(out)
    print("INTEGRATION.md written (%d bytes)" % len(out))
