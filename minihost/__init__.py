"""In-process AviSynth+ C-API stand-in and its Python driver: test/bench infrastructure."""
