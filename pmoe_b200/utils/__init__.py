"""Host-side helpers mirroring the reference's utils package (checkpoint io, gradient-norm check, freeze)."""
