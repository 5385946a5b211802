"""Mirror of the ``gsplat.cuda`` package path that MTGS imports from."""
