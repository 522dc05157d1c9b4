"""the Phantom dump reader that lands particle columns on the device.

`-m "not gpu"`: the record-structure parsing, on hand-built dumps (true Fortran length tags, both byte
orders, all four default-kind combinations, MPI blocks, sinks, inactive particles, mixed particle types)
with device='cpu', and -- when the reference is importable -- against `sarracen.read_phantom` on the same
files and on a dump written by `sarracen.write_phantom`.  `-m gpu`: columns streamed to the GPU render the
same image as the host frame."""
import numpy as np
import pytest

from oracle import ref_shim


def _rec(payload: bytes, order: str) -> bytes:
    tag = np.array([len(payload)], dtype=order + "i4").tobytes()
    return tag + payload + tag


def build_dump(arrays, sinks=None, def_int="i4", def_real="f8", order="<", mpi_blocks=1, header=None):
    """A Phantom dump as bytes.  `arrays` / `sinks`: {tag: ndarray}; the kind of each array is chosen from its
    dtype (default kinds first).  With mpi_blocks > 1 the particles are split evenly."""
    it, rt = np.dtype(order + def_int), np.dtype(order + def_real)
    kinds = [it, np.dtype("i1"), np.dtype(order + "i2"), np.dtype(order + "i4"), np.dtype(order + "i8"),
             rt, np.dtype(order + "f4"), np.dtype(order + "f8")]
    out = _rec(np.array([60769], it).tobytes() + np.array([60878], rt).tobytes() + np.array([60878], it).tobytes()
               + np.array([1], it).tobytes() + np.array([690706], it).tobytes(), order)
    out += _rec("FT:Phantom test dump".ljust(100).encode("ascii"), order)
    header = dict(header or {})
    header.setdefault("massoftype", np.float64(1e-6))
    per_kind = [dict() for _ in kinds]
    per_kind[0]["nblocks"] = mpi_blocks
    for k, v in header.items():
        per_kind[0 if isinstance(v, (int, np.integer)) else 5][k] = v
    for kd, vals in zip(kinds, per_kind):
        out += _rec(np.array([len(vals)], order + "i4").tobytes(), order)
        if vals:
            out += _rec("".join(k.split("#")[0].ljust(16) for k in vals).encode("ascii"), order)
            out += _rec(np.array(list(vals.values()), kd).tobytes(), order)

    def kind_of(a):
        dt = np.dtype(a.dtype.str.lstrip("<>=|"))
        for i, kd in enumerate(kinds):
            if np.dtype(kd.str.lstrip("<>=|")) == dt:
                return i
        raise ValueError(dt)

    blocks_per = 2 if sinks else 1
    out += _rec(np.array([blocks_per * mpi_blocks], order + "i4").tobytes(), order)
    n = len(next(iter(arrays.values())))
    bounds = [(n * j // mpi_blocks, n * (j + 1) // mpi_blocks) for j in range(mpi_blocks)]
    for lo, hi in bounds:
        groups = [{k: v[lo:hi] for k, v in arrays.items()}] + ([sinks] if sinks else [])
        for g in groups:
            nums = np.zeros(8, order + "i4")
            for a in g.values():
                nums[kind_of(a)] += 1
            out += _rec(np.array([len(next(iter(g.values())))], order + "i8").tobytes() + nums.tobytes(), order)
        for g in groups:
            for kd_i, kd in enumerate(kinds):
                for tag, a in g.items():
                    if kind_of(a) == kd_i:
                        out += _rec(tag.split("#")[0].ljust(16).encode("ascii"), order)
                        out += _rec(np.ascontiguousarray(a).astype(kd).tobytes(), order)
    return out


def _particles(n=1000, seed=3, real="f8"):
    rng = np.random.default_rng(seed)
    a = {k: rng.random(n).astype(real) for k in ("x", "y", "z", "vx", "vy", "vz")}
    a["h"] = rng.uniform(0.01, 0.05, n).astype("f4")
    a["h"][::17] = -a["h"][::17]                 # inactive (accreted) particles
    a["itype"] = np.ones(n, "i1")
    return a


@pytest.mark.parametrize("def_int,def_real,order", [("i4", "f8", "<"), ("i4", "f4", "<"), ("i8", "f8", "<"),
                                                    ("i8", "f4", ">"), ("i4", "f8", ">")])
def test_hand_built_dump_round_trip(tmp_path, def_int, def_real, order):
    from sarracen_b200.readers import read_phantom
    arrays = _particles(real=def_real)
    sinks = {"x": np.array([0.5, 0.25], def_real), "m": np.array([1.0, 2.0], def_real)}
    path = tmp_path / "dump"
    path.write_bytes(build_dump(arrays, sinks, def_int, def_real, order, mpi_blocks=2,
                                header={"hfact": np.float64(1.2), "time": np.float64(0.5), "npartoftype": 1000}))
    gas, snk = read_phantom(str(path), device="cpu")
    keep = arrays["h"] > 0
    assert len(gas) == int(keep.sum()) and gas.get_dim() == 3 and gas.hcol == "h"
    for k, v in arrays.items():
        np.testing.assert_array_equal(gas[k].to_numpy(), v[keep])
        assert gas.column_dtype(k) == v.dtype
    assert gas.params["mass"] == gas.params["massoftype"] == 1e-6 and gas.params["hfact"] == 1.2
    assert gas.params["nblocks"] == 2 and gas.params["def_real_dtype"] == np.dtype(def_real).type
    assert "mass" not in snk.params and list(snk["m"]) == [1.0, 2.0, 1.0, 2.0]   # sinks repeat per MPI block
    everything = read_phantom(str(path), device="cpu", separate_types=None, ignore_inactive=False)
    assert len(everything) == 1000 + 4 and np.isnan(everything["h"].to_numpy()[-1])


def test_mixed_particle_types_get_a_mass_column(tmp_path):
    from sarracen_b200.readers import read_phantom
    arrays = _particles(300)
    arrays["h"] = np.abs(arrays["h"])
    arrays["itype"][200:] = 7
    path = tmp_path / "dump"
    path.write_bytes(build_dump(arrays, header={"massoftype": np.float64(1e-6), "massoftype_7": np.float64(3e-8)}))
    sdf = read_phantom(str(path), device="cpu")
    assert sdf.mcol == "mass" and "mass" not in sdf.params
    np.testing.assert_array_equal(sdf["mass"].to_numpy(), np.where(arrays["itype"] == 7, 3e-8, 1e-6))
    gas, dust = read_phantom(str(path), device="cpu", separate_types="all")
    assert len(gas) == 200 and len(dust) == 100 and dust.params["mass"] == 3e-8


def test_not_a_phantom_file(tmp_path):
    from sarracen_b200.readers import read_phantom
    path = tmp_path / "junk"
    path.write_bytes(b"\x00" * 256)
    with pytest.raises(AssertionError):
        read_phantom(str(path), device="cpu")


@pytest.mark.skipif(not ref_shim.available(), reason="reference not present")
def test_against_the_reference_reader_and_writer(tmp_path):
    from sarracen_b200.readers import read_phantom
    sar = ref_shim.load_reference()
    arrays = _particles(500)
    sinks = {"x": np.array([0.5]), "y": np.array([0.1]), "z": np.array([0.2]), "m": np.array([1.0]),
             "h": np.array([0.3], "f4")}
    path = tmp_path / "dump"
    path.write_bytes(build_dump(arrays, sinks, header={"hfact": np.float64(1.2), "massoftype": np.float64(2e-6)}))
    ref_gas, ref_sinks = sar.read_phantom(str(path))
    gas, snk = read_phantom(str(path), device="cpu")
    assert list(gas.columns) == list(ref_gas.columns)
    for k in ref_gas.columns:
        np.testing.assert_array_equal(gas[k].to_numpy(), ref_gas[k].to_numpy())
    for k, v in ref_gas.params.items():
        assert gas.params[k] == v, k
    np.testing.assert_array_equal(snk["m"].to_numpy(), ref_sinks["m"].to_numpy())
    # a dump written by the reference's own writer
    out = tmp_path / "written"
    sar.write_phantom(str(out), ref_gas, ref_sinks)
    again = read_phantom(str(out), device="cpu")
    again = again[0] if isinstance(again, list) else again
    ref_again = sar.read_phantom(str(out))
    ref_again = ref_again[0] if isinstance(ref_again, list) else ref_again
    assert list(again.columns) == list(ref_again.columns)
    for k in ref_again.columns:
        np.testing.assert_array_equal(again[k].to_numpy(), ref_again[k].to_numpy())


@pytest.mark.gpu
def test_dump_to_device_to_image(tmp_path):
    """Columns streamed to the GPU by the reader render the same image as the same data given as a host frame,
    and match the oracle."""
    import sarracen_b200 as s
    from sarracen_b200 import interpolate as f
    from sarracen_b200.readers import read_phantom
    from oracle import OracleBackend, OracleKernel
    from conftest import assert_parity
    rng = np.random.default_rng(11)
    n = 200_000
    arrays = {k: rng.random(n) for k in ("x", "y", "z")}
    arrays["h"] = rng.uniform(0.004, 0.03, n).astype("f4")
    arrays["rho"] = rng.uniform(0.5, 2.0, n)
    path = tmp_path / "dump"
    path.write_bytes(build_dump(arrays, order=">", header={"hfact": np.float64(1.2),
                                                           "massoftype": np.float64(1.0 / n)}))
    sdf = read_phantom(str(path))
    assert all(t.is_cuda for t in sdf._device_cache.values()) and len(sdf) == n
    kw = dict(x_pixels=128, y_pixels=128, xlim=(0, 1), ylim=(0, 1), normalize=False)
    img = f.interpolate_3d_proj(sdf, "rho", **kw)
    host = s.ParticleFrame({k: v for k, v in arrays.items()}, params={"hfact": 1.2, "mass": 1.0 / n})
    assert_parity(img, f.interpolate_3d_proj(host, "rho", **kw), 1e-13, "device frame vs host frame")
    w = np.full(n, 1.0 / n)           # target rho, dens_weight: rho m / rho... 'rho' renders m (dens_weight False)
    ref = OracleBackend.interpolate_3d_projection(arrays["x"], arrays["y"], w, arrays["h"].astype(np.float64),
                                                  OracleKernel.column("cubic", 1000), 2.0, 128, 128, 0, 1, 0, 1, False)
    assert_parity(img, ref, 1e-10, "dump -> device -> image vs oracle")
