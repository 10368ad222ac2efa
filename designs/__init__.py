"""Design files of the elasticity path plus aliases of the schema/parser modules."""
