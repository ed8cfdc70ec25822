"""CPU oracle for the raster-vector overlay hot path of proj-roadsurf.

TEST INFRASTRUCTURE ONLY.  Nothing under ``proj_roadsurf_b200/`` may import,
call, link or execute anything in this package.  The only legitimate users are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- and there only as the checker / the timed
CPU baseline, never as the thing shipped.

Pinning status
--------------
* Rasterization, crop windows, masked extraction (``gdal_fill``, ``raster``):
  **parity unpinned**.  The arithmetic lives in GDAL 3.0.4 / rasterio 1.3.2 /
  rasterstats 0.17.0 (pins: /root/reference/requirements.txt:77,211,216), none
  of which is vendored in the reference or importable here, and the reference
  ships no test, golden vector or fixture for this path.  These modules restate
  the published algorithms and are anchored on the
  reference's call sites (scripts/functions/fct_misc.py:72-121,
  scripts/sandbox/add_tile_mask.py:112-113, scripts/functions/fct_rasters.py:162).
  Hand-derived known-answer tests live in tests/test_oracle_kat.py.
* Table logic (``stats``, ``vote``): pinned against the reference's own Python
  functions executed in the build container with the missing third-party
  modules stubbed (tests/golden/make_golden.py; fixtures in tests/golden/).
"""
