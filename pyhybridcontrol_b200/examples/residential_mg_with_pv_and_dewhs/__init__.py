"""The reference's example application (examples/residential_mg_with_pv_and_dewhs/): DEWH / PV / residential
demand / grid device models and the synthetic workload that stands in for its missing data pickles."""
