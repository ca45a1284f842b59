"""B200-native (sm_100a) transport-matrix assembly behind OceanTransportMatrixBuilder.jl's API.

The directory name carries a dot, so import it through the repository-root alias module:
`import otmb_b200` (see /otmb_b200.py).  The CUDA library (libotmb.so) is loaded on first
use and a missing library is an ImportError: there is no CPU fallback.
"""
from .api import (  # noqa: F401
    Context, Field, FaceFluxes, GridMetrics, GridTopology, Indices, OTMBError, TransportMatrices,
    bolus_GM_velocity, default_context, facefluxes, facefluxesfrommasstransport, getgridtopology,
    globalverticaldyadderivative, globalverticalfacetriadderivative, makegridmetrics, makeindices,
    spadd, sparse, transportmatrix, vertexpermutation, velocity2fluxes, fluxes2velocity, facefluxesfromvelocities,
    getarakawagrid, interpolateontodefaultCgrid, lump_and_spray, resident_matvec, facefluxes_GM, coarsen, dump_resident, load_dump,
)
from . import synthetic  # noqa: F401

__all__ = ["makegridmetrics", "makeindices", "facefluxesfrommasstransport", "facefluxesfromvelocities", "velocity2fluxes",
           "fluxes2velocity", "transportmatrix"]        # the reference's exports, src/OceanTransportMatrixBuilder.jl:31-36
