"""One rank of the multi-process sharding test (launched by tests/test_multi_rank.py through
torch.distributed.run with the gloo backend).  Each rank renders ITS shard of the font x GlyphBlock task
list with the dummy renderer (no GPU here); rank 0 checks that the shards are disjoint, that their union
is the whole job and that the gathered bytes equal a single-process run."""
import hashlib
import json
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (fixture paths only)
import versatiles_glyphs_rs_b200 as V  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    m.add_font_with_name("Noto Sans Regular", O.noto_paths()[:3])
    r = V.Renderer.new_dummy()
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, r, shard=rank, n_shards=world, threads=2)
    mine = {n: hashlib.sha1(d).hexdigest() for n, is_dir, d in w.entries() if not is_dir}
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, st.glyphs, st.blocks, st.cost_shard, st.cost_total))
    if rank == 0:
        full = V.Writer.new_memory()
        fst = m.render_glyphs(full, r)
        want = {n: hashlib.sha1(d).hexdigest() for n, is_dir, d in full.entries() if not is_dir}
        union = {}
        for files, *_ in gathered:
            assert not set(files) & set(union), "shards overlap"
            union.update(files)
        assert union == want, "union of shards != whole job"
        assert sum(g[1] for g in gathered) == fst.glyphs and sum(g[2] for g in gathered) == fst.blocks == 512
        # every rank computed the same cost tables and the same longest-processing-time-first assignment
        assert len({g[4] for g in gathered}) == 1 and sum(g[3] for g in gathered) == gathered[0][4] > 0
        print(json.dumps({"ok": True, "world": world, "files": len(union), "glyphs": fst.glyphs,
                          "per_rank_blocks": [g[2] for g in gathered], "per_rank_cost": [g[3] for g in gathered]}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
