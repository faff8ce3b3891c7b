"""CPU oracle for the final-product fields: numpy restatement of the array preparation in the reference's
``OutputSnowModelFinal`` (``/root/reference/source/utils.py:161-179``) fed the way ``main`` feeds it
(``NESOSIM.py:654``).  TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own function executed verbatim with a
recording stand-in for ``netCDF4`` (``tests/test_final_products.py::test_oracle_matches_reference_writer``)."""
import numpy as np


def final_fields(snowDepths, density, iceConcDays, precipDays, windDays, ice_conc_mask=0.5):
    """Returns the six float32 (T,ny,nx) fields the NetCDF file stores.  Inputs are not modified (the reference
    mutates ``iceConcDays`` in place, utils.py:167)."""
    with np.errstate(all="ignore"):
        snowVolT = snowDepths[:, 0] + snowDepths[:, 1]                     # NESOSIM.py:654
        snowDepthT = (snowDepths[:, 0] + snowDepths[:, 1]) / iceConcDays
        densityT = np.array(density, dtype=np.float64)
        iceConcT = np.array(iceConcDays, dtype=np.float64)
        if ice_conc_mask > 0:                                              # utils.py:161-167
            snowVolT[np.where(iceConcT < ice_conc_mask)] = np.nan
            snowDepthT[np.where(iceConcT < ice_conc_mask)] = np.nan
            densityT[np.where(iceConcT < ice_conc_mask)] = np.nan
            iceConcT[np.where(iceConcT < 0.15)] = np.nan
        f4 = lambda a: np.asarray(np.around(a, decimals=4), dtype=np.float32)   # utils.py:175-180 + the 'f4' variables
        return {"snow_volume": f4(snowVolT), "snow_depth": f4(snowDepthT), "snow_density": f4(densityT),
                "ice_concentration": f4(iceConcT), "precipitation": f4(precipDays), "wind_speed": f4(windDays)}
