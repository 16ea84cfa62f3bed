"""Pins the data-feed restatement against the reference itself (runs only where /root/reference exists) and writes
tests/golden/data_feed/: a tiny Megatron .bin/.idx corpus WRITTEN BY THE REFERENCE's builder plus the samples, indices
and rank assignment the reference's GPTDataset / BlendedDataset / MegatronBatchSampler produce from it.

    python oracle/pin_data_feed.py

Test infrastructure only.  The reference's helpers.cpp is compiled from where it lies into oracle/_ref/ (git-ignored);
its Python modules are imported read-only through a stub package chain (the real `dolomite_engine` package does not
import under this image's torch / transformers)."""
import importlib.util
import os
import subprocess
import sys
import sysconfig
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/dolomite_engine"
OUT = os.path.join(ROOT, "tests", "golden", "data_feed")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def compile_reference_helpers():
    import pybind11

    os.makedirs(REFDIR, exist_ok=True)
    so = os.path.join(REFDIR, "helpers" + sysconfig.get_config_var("EXT_SUFFIX"))
    src = os.path.join(REF, "data", "megatron", "utils", "helpers.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["g++", "-O3", "-shared", "-std=c++17", "-fPIC", f"-I{pybind11.get_include()}",
               f"-I{sysconfig.get_paths()['include']}", src, "-o", so]
        subprocess.run(cmd, check=True)
    sys.path.insert(0, REFDIR)
    import helpers  # noqa: F401

    return sys.modules["helpers"]


def import_reference_data_modules():
    """dolomite_engine.data.megatron.{indexed_dataset, gpt_dataset, blended_dataset, sampler, ...} with stubbed parents"""

    def pkg(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    pkg("dolomite_engine")
    utils = pkg("dolomite_engine.utils")
    utils.log_rank_0 = lambda *a, **k: None

    class PGM:
        @staticmethod
        def get_global_rank():
            return 0

        @staticmethod
        def get_data_parallel_world_size():
            return 1

    utils.ProcessGroupManager = PGM
    utils.run_rank_n = lambda f, *a, **k: f
    pkg("dolomite_engine.data", os.path.join(REF, "data"))
    pkg("dolomite_engine.data.megatron", os.path.join(REF, "data", "megatron"))

    def load(name, rel, is_pkg=False):
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REF, rel), submodule_search_locations=[os.path.dirname(os.path.join(REF, rel))] if is_pkg else None)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    # utils/__init__.py imports torch.utils.cpp_extension.load and `....utils.log_rank_0`: fine with the stubs
    load("dolomite_engine.data.megatron.utils", "data/megatron/utils/__init__.py", is_pkg=True)
    mods = {}
    for n in ("indexed_dataset", "blended_megatron_dataset_config", "megatron_dataset", "gpt_dataset", "blended_dataset",
              "sampler", "blended_megatron_dataset_builder"):
        mods[n] = load(f"dolomite_engine.data.megatron.{n}", f"data/megatron/{n}.py")
    return mods


def main():
    helpers = compile_reference_helpers()
    m = import_reference_data_modules()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)

    # ---- corpora written by the reference's builder (uint16 and int32 token ids, one empty document) ----
    corpora = {}
    for name, dtype, n_docs, vocab in (("corpus_a", np.uint16, 37, 5000), ("corpus_b", np.int32, 23, 70000)):
        prefix = os.path.join(OUT, name)
        b = m["indexed_dataset"].MMapIndexedDatasetBuilder(prefix + ".bin", dtype=dtype)
        docs = []
        import torch

        for d in range(n_docs):
            n = 0 if (name == "corpus_a" and d == 11) else int(rng.integers(1, 90))
            toks = rng.integers(0, vocab, size=n)
            docs.append(toks)
            b.add_item(torch.from_numpy(toks.astype(np.int64)))
            b.end_document()
        b.finalize(prefix + ".idx")
        corpora[name] = docs

    out = {}
    ids_a = m["indexed_dataset"].MMapIndexedDataset(os.path.join(OUT, "corpus_a"))
    ids_b = m["indexed_dataset"].MMapIndexedDataset(os.path.join(OUT, "corpus_b"))
    out["a_sequence_lengths"] = np.asarray(ids_a.sequence_lengths)
    out["a_document_indices"] = np.asarray(ids_a.document_indices)
    out["a_doc5"] = np.asarray(ids_a[5])
    out["b_doc3_slice"] = np.asarray(ids_b.get(3, offset=1, length=4))

    # ---- GPTDataset of the reference: several (split range, num_samples, seq) cases incl. multi-epoch + separate final epoch
    Cfg = m["blended_megatron_dataset_config"].GPTDatasetConfig
    Split = sys.modules["dolomite_engine.data.megatron.utils"].Split
    cases = [("a", ids_a, 0, 30, 25, 16, 1234), ("a", ids_a, 0, 37, 400, 8, 99), ("a", ids_a, 30, 37, 61, 12, 5),
             ("b", ids_b, 0, 23, 57, 32, 1234), ("b", ids_b, 2, 20, 300, 7, 3)]
    for ci, (tag, ids, lo, hi, num_samples, S, seed) in enumerate(cases):
        cache = os.path.join("/tmp", f"pin_data_feed_cache_{ci}")
        cfg = Cfg(is_built_on_rank=True, random_seed=seed, sequence_length=S, blend=[os.path.join(OUT, "corpus_" + tag)],
                  split="100,0,0", path_to_cache=cache, return_document_ids=False, fim_rate=0, fim_spm_rate=0.5)
        ds = m["gpt_dataset"].GPTDataset(ids, np.arange(lo, hi, dtype=np.int32), num_samples, Split.train, None, cfg, True)
        out[f"case{ci}_meta"] = np.asarray([lo, hi, num_samples, S, seed, len(ds)], dtype=np.int64)
        out[f"case{ci}_document_index"] = np.asarray(ds.document_index)
        out[f"case{ci}_sample_index"] = np.asarray(ds.sample_index)
        out[f"case{ci}_shuffle_index"] = np.asarray(ds.shuffle_index)
        take = list(range(min(len(ds), 40))) + [len(ds) - 1]
        out[f"case{ci}_take"] = np.asarray(take, dtype=np.int64)
        out[f"case{ci}_samples"] = np.stack([ds[i]["text"] for i in take])

    # ---- the index cache AS THE REFERENCE WRITES IT (gpt_dataset.py:265-330): relative prefix + relative cache directory, so that
    # the unique description (and its MD5 = the file names) does not depend on where the repository lives.  Committed under
    # tests/golden/data_feed/ref_index_cache/: the product must find these files (cache hit) and must name its own files alike.
    import shutil

    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        shutil.rmtree("ref_index_cache", ignore_errors=True)
        ids_rel = m["indexed_dataset"].MMapIndexedDataset("corpus_a")
        cfg = Cfg(is_built_on_rank=True, random_seed=77, sequence_length=16, blend=["corpus_a"], split="90,10,0",
                  path_to_cache="ref_index_cache", return_document_ids=False, fim_rate=0, fim_spm_rate=0.5)
        ds = m["gpt_dataset"].GPTDataset(ids_rel, np.arange(0, 33, dtype=np.int32), 230, Split.train, None, cfg, True)
        out["cache_description"] = np.asarray(ds.unique_description)
        out["cache_hash"] = np.asarray(ds.unique_description_hash)
        out["cache_files"] = np.asarray(sorted(os.listdir("ref_index_cache")))
        out["cache_len"] = np.asarray(len(ds))
        out["cache_samples"] = np.stack([ds[i]["text"] for i in range(0, len(ds), 7)])
        # the validation split of the same configuration: only the description differs (index_split, num_samples)
        dv = m["gpt_dataset"].GPTDataset(ids_rel, np.arange(33, 37, dtype=np.int32), 9, Split.valid, None, cfg, True)
        out["cache_hash_valid"] = np.asarray(dv.unique_description_hash)
    finally:
        os.chdir(cwd)

    # ---- the reference's BUILDER (blended_megatron_dataset_builder.py): option 2 (weighted blend cut by `split`) and option 3
    # (a blend per split): lengths of what it returns and the first samples of every split
    Builder = m["blended_megatron_dataset_builder"].BlendedMegatronDatasetBuilder
    pa, pb = os.path.join(OUT, "corpus_a"), os.path.join(OUT, "corpus_b")
    builder_cases = {
        "opt2": (dict(blend=["1", pa, "3", pb], split="90,10,0"), [40, 8, 0], 8, 1234),
        "opt3": (dict(blend_per_split=[["2", pa, "1", pb], [pb], None]), [50, 6, 0], 8, 5),
    }
    for tag, (kw, sizes, S, seed) in builder_cases.items():
        cfg = Cfg(is_built_on_rank=True, random_seed=seed, sequence_length=S, path_to_cache=os.path.join("/tmp", f"pin_builder_{tag}"),
                  return_document_ids=False, fim_rate=0, fim_spm_rate=0.5, **kw)
        splits = Builder(m["gpt_dataset"].GPTDataset, sizes, cfg, None).build()
        out[f"builder_{tag}_lens"] = np.asarray([-1 if d is None else len(d) for d in splits], dtype=np.int64)
        for i, d in enumerate(splits):
            if d is None:
                continue
            n = min(len(d), 24)
            out[f"builder_{tag}_split{i}_samples"] = np.stack([d[j]["text"] for j in range(n)])
            if hasattr(d, "dataset_index"):
                out[f"builder_{tag}_split{i}_dataset_index"] = np.asarray(d.dataset_index)
                out[f"builder_{tag}_split{i}_dataset_sample_index"] = np.asarray(d.dataset_sample_index)

    # ---- blending index + sampler straight from the reference helpers / class ----
    for bi, (w, size) in enumerate([([0.3, 0.7], 101), ([0.5, 0.25, 0.25], 64), ([1.0], 9), ([0.2, 0.2, 0.6], 1000)]):
        di = np.zeros(size, dtype=np.int16)
        dsi = np.zeros(size, dtype=np.int64)
        helpers.build_blending_indices(di, dsi, np.asarray(w, dtype=np.float64), len(w), size, False)
        out[f"blend{bi}_weights"] = np.asarray(w)
        out[f"blend{bi}_dataset_index"] = di
        out[f"blend{bi}_dataset_sample_index"] = dsi
    for si, (total, consumed, mbs, world, drop_last) in enumerate([(100, 0, 4, 2, True), (103, 16, 3, 4, True), (50, 10, 4, 3, False)]):
        for rank in range(world):
            s = m["sampler"].MegatronBatchSampler(total, consumed, mbs, world, rank, drop_last)
            rows = list(s)
            flat = np.asarray([x for r in rows for x in r] or [-1], dtype=np.int64)
            lens = np.asarray([len(r) for r in rows], dtype=np.int64)
            out[f"sampler{si}_rank{rank}_flat"] = flat
            out[f"sampler{si}_rank{rank}_lens"] = lens
        out[f"sampler{si}_meta"] = np.asarray([total, consumed, mbs, world, int(drop_last)], dtype=np.int64)
    # raw sample-index helper on adversarial sizes (zero-length documents, exact boundary hits)
    sizes = np.asarray([5, 0, 3, 8, 0, 0, 1, 16, 2, 7], dtype=np.int32)
    doc_idx32 = np.asarray([3, 1, 0, 9, 4, 7, 2, 8, 6, 5, 0, 3, 7, 9, 1], dtype=np.int32)
    tpe = int(sizes.sum())
    for S in (2, 3, 4, 8):
        ne = 1
        out[f"raw_S{S}_i32"] = helpers.build_sample_idx_int32(sizes, doc_idx32[:10], S, ne, tpe)
        out[f"raw_S{S}_i64"] = helpers.build_sample_idx_int64(sizes, doc_idx32[:10].astype(np.int64), S, ne, tpe)
    out["raw_sizes"], out["raw_doc_idx"] = sizes, doc_idx32[:10]
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **out)
    print("wrote", OUT, {k: v.shape for k, v in list(out.items())[:6]}, "...", len(out), "arrays")


if __name__ == "__main__":
    main()
