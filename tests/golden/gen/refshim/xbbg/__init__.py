class blp: pass
