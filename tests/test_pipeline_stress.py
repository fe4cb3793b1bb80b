"""FontManager::render_glyphs under many thread counts (dummy renderer, no GPU): finished batches are encoded part by part
by whichever workers are free, merged submissions come back together — whatever the interleaving, every call must write
exactly the files of a single-threaded call."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O  # noqa: E402  (fixture paths)
import versatiles_glyphs_rs_b200 as V  # noqa: E402


def _run(manager, renderer, threads):
    w = V.Writer.new_memory()
    st = manager.render_glyphs(w, renderer, threads=threads)
    h = hashlib.sha1()
    for name, is_dir, data in sorted(w.entries()):
        h.update(name.encode())
        h.update(hashlib.sha1(data).digest())
    return st, h.hexdigest()


def test_every_interleaving_writes_the_same_files():
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Noto Sans Regular", O.noto_paths())
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    r = V.Renderer.new_dummy()
    st0, d0 = _run(m, r, 1)
    assert st0.glyphs == 6480 + 1686 and st0.blocks == 512
    for threads in (2, 3, 4, 8, 16, 33):
        for k in range(25):
            st, d = _run(m, r, threads)
            assert (st.glyphs, st.blocks, st.bitmaps) == (st0.glyphs, st0.blocks, st0.bitmaps), (threads, k)
            assert d == d0, (threads, k)


def test_directory_sink_written_by_several_workers_at_once(tmp_path):
    """The directory sink needs no lock between files: workers write theirs side by side, the tree equals the
    single-threaded one."""
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Noto Sans Regular", O.noto_paths())
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    r = V.Renderer.new_dummy()

    def tree(threads, name):
        d = tmp_path / name
        d.mkdir()
        w = V.Writer.new_file(str(d))
        st = m.render_glyphs(w, r, threads=threads)
        h, n = hashlib.sha1(), 0
        for root, _, files in sorted(os.walk(d)):
            for f in sorted(files):
                p = os.path.join(root, f)
                h.update(os.path.relpath(p, d).encode())
                h.update(open(p, "rb").read())
                n += 1
        return h.hexdigest(), n, st.blocks

    want = tree(1, "one")
    assert want[1] == 512 == want[2]
    for k, threads in enumerate((4, 8, 16, 8)):
        assert tree(threads, f"t{k}") == want
