# box.py -- input script for the reference's chkn-prep (src/chicken/chkn_prep.py): the 3D ideal-air box of
# BASELINE.json configs[2] in chicken's own terms -- uniform supersonic flow through a unit cube, 2 x 2 x 2 blocks,
# inflow west, outflow east, slip walls elsewhere, AUSMDV, second-order reconstruction.  Run once in the build
# container by oracle/chicken/make_template.sh; its config.json is the template oracle/chicken/make_job.py scales.
config.title = "3D ideal-air box, 2x2x2 blocks"
N = 16
vol0 = TFIVolume(p000=Vector3(0.0, 0.0, 0.0), p100=Vector3(1.0, 0.0, 0.0),
                 p110=Vector3(1.0, 1.0, 0.0), p010=Vector3(0.0, 1.0, 0.0),
                 p001=Vector3(0.0, 0.0, 1.0), p101=Vector3(1.0, 0.0, 1.0),
                 p111=Vector3(1.0, 1.0, 1.0), p011=Vector3(0.0, 1.0, 1.0))
grd0 = StructuredGrid(pvolume=vol0, niv=N+1, njv=N+1, nkv=N+1)
inflow = FlowState(p=95.84e3, T=1103.0, velx=1000.0)
makeFBArray(ni=2, nj=2, nk=2, grid=grd0, initialState=inflow,
            bcs={'iminus': InflowBC(inflow), 'iplus': OutflowBC()})
config.max_time = 1.0
config.max_step = 20
config.print_count = 5
config.flux_calc = "ausmdv"
add_cfl_value(0.0, 0.5)
add_dt_plot(0.0, 1.0)
