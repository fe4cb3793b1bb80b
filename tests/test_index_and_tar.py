"""CPU tests (no GPU): the sink side of the path against the reference's own
known-answer tests — parse_font_name vectors (tests/golden/, extracted from
src/font/parse_font_name.rs), FontMetadata (src/font/metadata.rs:134-153), encode_codeblocks and
font_families.json / index.json (src/font/index_files.rs:139-228, src/font/manager.rs:225-262), and
the ustar writer (src/writer/tar.rs:180-300)."""
import io
import json
import os
import tarfile

import pytest

import oracle_lib as O
import versatiles_glyphs_rs_b200 as V

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOTO_REGULAR = os.path.join(O.NOTO_DIR, "Noto Sans - Regular.ttf")


def test_parse_font_name_reference_vectors():
    doc = json.load(open(os.path.join(GOLDEN, "parse_font_name_vectors.json")))
    vectors = doc["vectors"]
    assert len(vectors) == 243
    for v in vectors:
        got = V.parse_font_name(v["family"], v["ps_name"])
        assert list(got) == v["expect"], (v, got)


def test_parse_font_name_doc_example():
    # parse_font_name.rs:203-212
    assert V.parse_font_name("Open Sans SemiCondensed Light", "OpenSansSemiCondensed-LightItalic") == (
        "Open Sans", "italic", 300, "semi-condensed")
    assert V.parse_font_name("Foo Extra Condensed", "Foo") == ("Foo", "normal", 400, "extra-condensed")


def test_font_metadata_goldens():
    # metadata.rs:134-153
    m = V.FontFileEntry(path=O.FIRA).metadata
    assert (m["family"], m["generated_name"]) == ("Fira Sans", "Fira Sans Regular")
    m = V.FontFileEntry(path=NOTO_REGULAR).metadata
    assert (m["family"], m["generated_name"]) == ("Noto Sans", "Noto Sans Regular")
    assert (m["style"], m["weight"], m["width"]) == ("normal", 400, "normal")


def test_encode_codeblocks_goldens():
    # index_files.rs:209-227
    assert V.encode_codeblocks([]) == ""
    assert V.encode_codeblocks([0xA3]) == "A"
    assert V.encode_codeblocks([0x0, 0x1, 0x2, 0xF, 0x10]) == "0-1"
    assert V.encode_codeblocks([0x0, 0x2, 0x1F, 0x40, 0xA0]) == "0-1,4,A"


FAMILIES_GOLDEN = [  # index_files.rs:170-205
    "[",
    "  {",
    "    \"name\": \"Fira Sans\",",
    "    \"faces\": [",
    "      {",
    "        \"id\": \"fira_sans_regular\",",
    "        \"style\": \"normal\",",
    "        \"weight\": 400,",
    "        \"width\": \"normal\",",
    "        \"codeblocks\": \"0,2-7,A-2E,30-52,E3,1D4,1D6-1D7,1D9,1DB-1DC,1E0-204,207-208,20A-20B,210-212,215,219,21E,220-222,224,226,22C,232,23C,25A,25C,2C6-2C7,A78,A7A-A7B,AB5,FB0,FEF\"",
    "      }",
    "    ]",
    "  },",
    "  {",
    "    \"name\": \"Noto Sans\",",
    "    \"faces\": [",
    "      {",
    "        \"id\": \"noto_sans_regular\",",
    "        \"style\": \"normal\",",
    "        \"weight\": 400,",
    "        \"width\": \"normal\",",
    "        \"codeblocks\": \"0,2-7,A-52,90-97,10F,1AB-1AC,1C8,1D0-20C,20F-215,218,221,25C,2C6-2C7,2DE-2E5,A64-A69,A70-A7D,A7F,A8F,A92,AB3-AB6,FB0,FE0,FE2,FEF,FFF,1078-107B,1DF0-1DF1\"",
    "      }",
    "    ]",
    "  }",
    "]",
]


def _two_font_manager():
    m = V.FontManager(parallel=False)
    m.add_paths([O.FIRA, NOTO_REGULAR])  # add_path derives the id with parse_font_name (manager.rs:39-53)
    assert m.font_ids() == ["fira_sans_regular", "noto_sans_regular"]
    return m


def test_families_and_index_json_goldens():
    m = _two_font_manager()
    w = V.Writer.new_memory()
    m.write_families_json(w)
    m.write_index_json(w)
    files = {name: data for name, is_dir, data in w.entries()}
    assert sorted(files) == ["font_families.json", "index.json"]
    assert files["font_families.json"].decode().split("\n") == FAMILIES_GOLDEN
    # index_files.rs:152-160
    assert files["index.json"].decode().split("\n") == ["[", "  \"fira_sans_regular\",", "  \"noto_sans_regular\"", "]"]
    # manager.rs:239-242 compares the dummy writer's whitespace-collapsed record
    flat = "font_families.json: " + files["font_families.json"].decode().replace("\n", "").replace("  ", "")
    assert flat[:64] == "font_families.json: [{\"name\": \"Fira Sans\",\"faces\": [{\"id\": \"fira"


def test_families_json_groups_faces_by_family():
    m = V.FontManager(parallel=False)
    names = [n for n in sorted(os.listdir(O.NOTO_DIR)) if n.endswith(".ttf")][:4]
    for n in names:
        m.add_font_with_name(n[:-4], [os.path.join(O.NOTO_DIR, n)])
    w = V.Writer.new_memory()
    m.write_families_json(w)
    doc = json.loads(w.entries()[0][2])
    assert [f["name"] for f in doc] == sorted(f["name"] for f in doc)
    assert sum(len(f["faces"]) for f in doc) == len(names)
    for fam in doc:
        for face in fam["faces"]:
            assert set(face) == {"id", "style", "weight", "width", "codeblocks"}


def _until_nul(b):
    return b.split(b"\0", 1)[0].decode()


def test_tar_long_filename_errors():
    w = V.Writer.new_tar_memory()
    with pytest.raises(V.B200Error, match="tar header field overflow"):
        w.write_file("a" * 101, b"x")


def test_tar_write_file():
    w = V.Writer.new_tar_memory()
    w.write_file("testfile.txt", b"hello tar")
    w.finish()
    out = w.tar_bytes()
    assert len(out) == 2048
    assert _until_nul(out[0:100]) == "testfile.txt"
    assert out[156:157] == b"0"
    assert out[512:521] == b"hello tar"
    assert out[521:1024] == bytes(503)
    w.finish()  # idempotent (writer/mod.rs:67-73)
    assert len(w.tar_bytes()) == 2048


def test_tar_write_directory():
    w = V.Writer.new_tar_memory()
    w.write_directory("testdir/")
    w.finish()
    out = w.tar_bytes()
    assert len(out) == 1536
    assert _until_nul(out[0:100]) == "testdir/"
    assert out[156:157] == b"5"
    assert out[512:] == bytes(1024)
    with pytest.raises(V.B200Error, match="must end with a slash"):
        V.Writer.new_tar_memory().write_directory("nodash")


def test_tar_multiple_files_and_finish():
    w = V.Writer.new_tar_memory()
    w.write_file("file1.txt", b"foo")
    w.write_file("file2.txt", b"barbaz")
    w.finish()
    out = w.tar_bytes()
    assert len(out) == 3072
    assert _until_nul(out[0:100]) == "file1.txt"
    assert _until_nul(out[1024:1124]) == "file2.txt"
    assert out[512:515] == b"foo" and out[1536:1542] == b"barbaz"


def test_tar_real_decoder():
    """tar.rs:259-289 with Python's tarfile in place of the `tar` crate."""
    w = V.Writer.new_tar_memory()
    w.write_file("file1.txt", b"content 1")
    w.write_directory("folder/")
    w.write_file("file2.txt", b"content 2")
    w.write_file("folder/file3.txt", b"content 3")
    w.finish()
    out = w.tar_bytes()
    tf = tarfile.open(fileobj=io.BytesIO(out))
    members = tf.getmembers()
    got = [(("Directory" if m.isdir() else "Regular"), m.name + ("/" if m.isdir() else ""), m.offset, m.offset_data, m.size)
           for m in members]
    assert got == [
        ("Regular", "file1.txt", 0, 512, 9),
        ("Directory", "folder/", 1024, 1536, 0),
        ("Regular", "file2.txt", 1536, 2048, 9),
        ("Regular", "folder/file3.txt", 2560, 3072, 9),
    ]
    assert [tf.extractfile(m).read() for m in members if m.isfile()] == [b"content 1", b"content 2", b"content 3"]
    for m in members:
        assert (m.mode, m.uid, m.gid) == (0o755 if m.isdir() else 0o644, 0, 0)


def test_render_glyphs_into_tar_matches_directory_sink(tmp_path):
    """The whole sink: dummy renderer -> tar file, every member equal to the in-memory writer's entry."""
    m = _two_font_manager()
    r = V.Renderer.new_dummy()
    mem = V.Writer.new_memory()
    m.render_glyphs(mem, r)
    m.write_index_json(mem)
    m.write_families_json(mem)
    path = str(tmp_path / "glyphs.tar")
    tw = V.Writer.new_tar(path)
    m.render_glyphs(tw, r)
    m.write_index_json(tw)
    m.write_families_json(tw)
    tw.finish()
    want = {name: data for name, is_dir, data in mem.entries() if not is_dir}
    tf = tarfile.open(path)
    got = {mm.name: tf.extractfile(mm).read() for mm in tf.getmembers() if mm.isfile()}
    assert got == want
    assert sum(1 for n in got if n.endswith(".pbf")) == 512  # manager.rs:199: all 256 ranges per font


def test_writer_failure_mid_pipeline_is_an_error_not_a_hang():
    """A sink that starts failing after the first flush (/dev/full): render_glyphs reports the error from every
    worker configuration, leaves no thread stuck, and the manager renders normally afterwards."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    m = V.FontManager(parallel=True)
    m.add_font_with_name("Fira Sans - Regular", [O.FIRA])
    r = V.Renderer.new_dummy()
    for threads in (1, 3, 0):
        with pytest.raises(V.B200Error, match="writing tar"):
            m.render_glyphs(V.Writer.new_tar("/dev/full"), r, threads=threads)
    w = V.Writer.new_memory()
    st = m.render_glyphs(w, r)
    assert st.blocks == 256 and st.glyphs == 1686


# ---- `scan` of the recurse command (commands/recurse.rs:104-133 and its tests :150-330) -------------------
def test_scan_testdata_like_the_reference():
    """recurse.rs test_scan: without a fonts.json every file is added by path, and parse_font_name strips the script
    word, so all Noto Sans files merge into noto_sans_regular (20 here: JP / KR / SC are not in this checkout)."""
    m = V.FontManager(parallel=False)
    m.scan(O.TESTDATA)
    assert m.font_ids() == ["fira_sans_regular", "noto_sans_regular"]
    assert sorted(m.font_file_names("noto_sans_regular")) == [
        "Noto Sans", "Noto Sans Arabic", "Noto Sans Armenian", "Noto Sans Balinese", "Noto Sans Bengali",
        "Noto Sans Devanagari", "Noto Sans Ethiopic", "Noto Sans Georgian", "Noto Sans Gujarati", "Noto Sans Gurmukhi",
        "Noto Sans Hebrew", "Noto Sans Javanese", "Noto Sans Kannada", "Noto Sans Khmer", "Noto Sans Lao",
        "Noto Sans Myanmar", "Noto Sans Oriya", "Noto Sans Sinhala", "Noto Sans Tamil", "Noto Sans Thai"]
    assert m.font_file_names("fira_sans_regular") == ["Fira Sans"]
    # entries are visited in byte-wise name order: "Noto Sans - Regular.ttf" first, so it owns the shared code points
    assert m.font_file_names("noto_sans_regular")[0] == "Noto Sans"


def test_scan_fonts_json_manifest_and_non_font_files(tmp_path):
    import shutil

    # recurse.rs test_run_with_fonts_json_manifest + test_scan_skips_non_font_files
    d = tmp_path / "input"
    d.mkdir()
    shutil.copy(O.FIRA, d / "font.ttf")
    shutil.copy(O.FIRA, d / "ignored because of the manifest.ttf")
    (d / "fonts.json").write_text('[{"name": "Custom Merged Sans", "sources": ["font.ttf"], "comment": {"x": [1, 2]}}]')
    plain = tmp_path / "plain"
    (plain / "sub").mkdir(parents=True)
    (plain / "README.txt").write_text("this is not a font")
    (plain / "UPPER.TTF").write_bytes(open(O.FIRA, "rb").read())  # the extension test is case-sensitive
    shutil.copy(O.FIRA, plain / "sub" / "deep.otf")
    m = V.FontManager(parallel=False)
    m.scan(str(d))
    m.scan(str(plain))
    m.scan(str(tmp_path / "does not exist"))
    assert m.font_ids() == ["custom_merged_sans", "fira_sans_regular"]
    assert m.font_file_names("custom_merged_sans") == ["Fira Sans"]
    w = V.Writer.new_file(str(tmp_path / "glyphs"))
    m.render_glyphs(w, V.Renderer.new_dummy())
    m.write_index_json(w)
    m.write_families_json(w)
    out = tmp_path / "glyphs"
    assert (out / "custom_merged_sans" / "0-255.pbf").is_file() and (out / "fira_sans_regular" / "65280-65535.pbf").is_file()
    assert "custom_merged_sans" in (out / "index.json").read_text() and (out / "font_families.json").is_file()
    # a broken manifest is an error, not a silent skip
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "fonts.json").write_text('[{"name": "No sources"}]')
    with pytest.raises(V.B200Error, match="fonts.json"):
        V.FontManager(parallel=False).scan(str(bad))


def test_name_to_id_follows_the_reference_rules():
    """manager.rs:141-147: Unicode to_lowercase, runs of [-_\\s] (regex-lite: ASCII \\s) collapse to one '_', Unicode trim."""
    # the reference's own examples (manager.rs tests) plus non-ASCII names
    for name, want in [
        ("Noto Sans - Regular", "noto_sans_regular"), ("Fira  Sans_Bold", "fira_sans_bold"), ("  x y\t", "x_y"),
        ("ÄRIAL Bold", "ärial_bold"), ("ΑΒΓ_Δ-Ω", "αβγ_δ_ω"), ("Привет  Мир", "привет_мир"),
        ("ÄRIAL\u00a0Bold", "ärial\u00a0bold"),  # NBSP is not matched by regex-lite's ASCII \s ...
        ("\u00a0 Foo \u3000", "foo"),             # ... but str::trim removes Unicode white space at both ends
        ("ŁÓDŹ Ÿ", "łódź_ÿ"),
    ]:
        assert V.name_to_id(name) == want, (name, V.name_to_id(name))


def test_index_json_escapes_font_ids():
    """serde_json (manager.rs:128-131) escapes quotes, backslashes and control characters: the file stays valid JSON."""
    m = V.FontManager(parallel=False)
    m.add_font_bytes_with_name('Foo "Bar"\\ \x01', open(O.FIRA, "rb").read())
    w = V.Writer.new_memory()
    m.write_index_json(w)
    text = [d for n, is_dir, d in w.entries() if n == "index.json"][0].decode()
    assert json.loads(text) == ['foo_"bar"\\_\x01']
    assert '\\u0001' in text and '\\"' in text


def test_units_per_em_out_of_range_is_a_parse_error():
    """ttf-parser rejects a head table whose unitsPerEm is outside 16..=16384 (file_entry.rs:48: "Could not parse font data")."""
    import struct

    data = bytearray(open(O.FIRA, "rb").read())
    n = struct.unpack(">H", data[4:6])[0]
    head = next(struct.unpack(">I", data[12 + 16 * i + 8:12 + 16 * i + 12])[0] for i in range(n) if data[12 + 16 * i:12 + 16 * i + 4] == b"head")
    for upm in (0, 15, 16385):
        bad = bytearray(data)
        bad[head + 18:head + 20] = struct.pack(">H", upm)
        with pytest.raises(V.B200Error, match="Could not parse font data"):
            V.FontFileEntry(data=bytes(bad))
    ok = bytearray(data)
    ok[head + 18:head + 20] = struct.pack(">H", 16)
    assert V.FontFileEntry(data=bytes(ok)).units_per_em == 16


def test_tar_finish_reports_close_errors(tmp_path):
    """A tar sink whose flush / close fails makes finish() fail (BufWriter::flush, writer/tar.rs:133-137)."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    w = V.Writer.new_tar("/dev/full")
    w.write_file("a.pbf", b"x" * 100)  # buffered by stdio: the error surfaces at flush
    with pytest.raises(V.B200Error):
        w.finish()
    good = V.Writer.new_tar(str(tmp_path / "ok.tar"))
    good.write_file("a.pbf", b"x" * 100)
    good.finish()
    assert tarfile.open(str(tmp_path / "ok.tar")).getnames() == ["a.pbf"]
